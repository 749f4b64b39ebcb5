"""CPU model of the fused Sobel-3 normal-map kernel's fp32 screening (csrc/normalmap.cu: normal_rgb_packed).

The kernel takes floor(fp32 value) whenever the fp32 value is farther than 2^-10 from an integer and re-runs the
reference's fp64 sequence otherwise.  This model repeats the fp32 arithmetic in numpy (with rsqrt perturbed by +-2 ulp,
the documented bound of rsqrt.approx) and checks, against the oracle, that every pixel the kernel would accept without
the fp64 detour already has the reference's bytes — i.e. that the safety band really covers the fp32 error."""
import numpy as np
import pytest

from synth import noise_depth_u16, synth_depth_u16

F = np.float32


def _fp32_screen(ax, ay, ulps):
    fx = ax.astype(F) * F(0.00390625)
    fy = ay.astype(F) * F(0.00390625)
    s = (fx.astype(np.float64) * fx + (fy.astype(np.float64) * fy + 1.0).astype(F)).astype(F)      # two fused multiply-adds
    rs = (1.0 / np.sqrt(s.astype(np.float64))).astype(F)
    rs = np.nextafter(rs, F(np.inf) if ulps > 0 else F(-np.inf)) if ulps else rs
    if abs(ulps) == 2:
        rs = np.nextafter(rs, F(np.inf) if ulps > 0 else F(-np.inf))
    r128 = rs * F(128)
    out = []
    for v in ((fx.astype(np.float64) * r128 + 128.0).astype(F), (fy.astype(np.float64) * r128 + 128.0).astype(F), (r128 + F(128)).astype(F)):
        t = (v + F(8388608.0)).astype(F)
        d = v - (t - F(8388608.0))
        safe = np.abs(d) > F(0.0009765625)
        q = np.minimum((t.view(np.int32) & 0x1ff) - (d < 0), 255)
        out.append((q, safe))
    return out


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("ulps", [-2, 0, 2])
def test_fp32_screen_never_accepts_a_wrong_byte(seed, ulps):
    from oracle import normalmap as onm
    h = w = 768
    for dep, inv in ((synth_depth_u16(h, w, seed), False), (noise_depth_u16(h, w, seed + 10), True),
                     ((noise_depth_u16(h, w, seed + 20) >> 9).astype(np.uint16), False)):      # small gradients: many near-ties
        want = onm.create_normalmap(dep, None, 3, None, inv, return_array=True)
        d = np.pad(dep.astype(np.int64), 1, mode="reflect")
        gx = (d[:-2, 2:] - d[:-2, :-2]) + 2 * (d[1:-1, 2:] - d[1:-1, :-2]) + (d[2:, 2:] - d[2:, :-2])
        gy = (d[2:, :-2] - d[:-2, :-2]) + 2 * (d[2:, 1:-1] - d[:-2, 1:-1]) + (d[2:, 2:] - d[:-2, 2:])
        sgn = 1 if inv else -1
        ax, ay = sgn * gx, -sgn * gy
        (qx, sx), (qy, sy), (qz, sz) = _fp32_screen(ax, ay, ulps)
        # exact special cases of the kernel
        qx = np.where(ax == 0, 128, qx); sx = sx | (ax == 0)
        qy = np.where(ay == 0, 128, qy); sy = sy | (ay == 0)
        both = (ax == 0) & (ay == 0)
        qz = np.where(both, 255, qz); sz = sz | both
        for c, (q, s) in enumerate(((qx, sx), (qy, sy), (qz, sz))):
            wrong = s & (q != want[..., c])
            assert not wrong.any(), (seed, ulps, inv, c, int(wrong.sum()))
        detour = ~(sx & sy & sz)
        assert detour.mean() < 0.05          # the fp64 detour stays rare (it is ~0.6% on natural depth maps)
