"""GPU: the depth networks on the BASELINE.json configurations (not only the small structural ones):
  C3 core  Depth-Anything-V2 ViT-L @518x518, batch > 1
  C2       dpt_beit_large_512 @512x512 at the bench batch (32) — compares images spread over the batch, because tile
           scheduling / L2 grouping of the persistent GEMMs depends on the batch
  C4 core  dpt_beit_large_384 (the ZoeDepth-NK trunk) incl. a non-square net (generic relative-position mode)
Bar: north_star's 1e-3 on depth = max |d_gpu - d_oracle| / (max - min of the oracle), with the reference's own precision
policy as the yardstick: the same network evaluated the way the reference runs it on a GPU (`model.half()`,
src/depthmap_generation.py:268-275 — fp16 weights AND fp16 activations / residual stream) is measured in the same test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_MAX, TOL_MEAN = 1e-3, 3e-4


class _HalfView:
    """state_dict view whose `.float()` hands back the fp16 CUDA copy: runs the oracle's functional network in the
    reference's GPU precision (everything fp16) without touching the oracle code."""

    def __init__(self, sd, dev):
        import torch
        self.d = {k: _HalfTensor(v.to(dev, torch.float16)) for k, v in sd.items()}

    def __getitem__(self, k):
        return self.d[k]

    def get(self, k, default=None):
        return self.d.get(k, default)


class _HalfTensor:
    def __init__(self, t):
        self.t = t

    def float(self):
        return self.t


def _err(got, want):
    rng = float(want.max() - want.min())
    return float(np.abs(got - want).max()) / rng, float(np.abs(got - want).mean()) / rng


def test_dav2_vitl_518(cuda_device):
    import torch
    import torch.nn.functional as F
    from depthmap_b200.depthmap_generation import DepthAnythingV2Engine
    from oracle import dav2 as odav2
    from oracle import synth_weights
    from synth import synth_rgb
    sd = synth_weights.make_dav2_state_dict('vitl', seed=1)
    eng = DepthAnythingV2Engine(sd, 'vitl', cuda_device)
    imgs = [synth_rgb(518, 518, 50 + s) for s in range(3)]
    got = eng.forward_batch(torch.from_numpy(np.stack(imgs)).to(cuda_device), 518).cpu().numpy()
    hv = _HalfView(sd, cuda_device)
    for i in (0, 2):
        want, _ = odav2.get_raw_prediction(imgs[i], sd, 'vitl', 518)
        assert want.max() - want.min() > 0.1
        mx, mean = _err(got[i], want)
        x, (h, w) = odav2.preprocess(imgs[i], 518)
        with torch.no_grad():
            ref16 = odav2.forward(hv, x.to(cuda_device, torch.float16), 'vitl').float()
            ref16 = F.interpolate(ref16[:, None], (h, w), mode="bilinear", align_corners=True)[0, 0].cpu().numpy()
        rmx, rmean = _err(ref16, want)
        print(f"dav2 vitl 518 img{i}: ours max {mx:.3e} mean {mean:.3e} | reference-style fp16 max {rmx:.3e} mean {rmean:.3e}")
        assert mx < TOL_MAX and mean < TOL_MEAN, (i, mx, mean, rmx, rmean)


def test_beit_large_512_batch32(cuda_device):
    import torch
    from depthmap_b200.depthmap_generation import DptBeitEngine
    from oracle import beit_dpt, synth_weights
    from synth import synth_rgb
    sd = synth_weights.make_beit_dpt_state_dict('beitl16_512', seed=3)
    eng = DptBeitEngine(sd, 'beitl16_512', cuda_device)
    uniq = [synth_rgb(512, 512, 70 + s) for s in range(3)]
    order = [0, 1, 2] * 10 + [1, 2]                        # 32 images, three distinct ones spread over the batch
    got = eng.forward_batch(torch.from_numpy(np.stack([uniq[k] for k in order])).to(cuda_device), 512, 512).cpu().numpy()
    wants = {}
    for pos in (0, 13, 31):
        k = order[pos]
        if k not in wants:
            wants[k] = beit_dpt.get_raw_prediction(uniq[k], sd, 'beitl16_512', 512, 512)[0]
        mx, mean = _err(got[pos], wants[k])
        print(f"beit512 B=32 img{pos}: max {mx:.3e} mean {mean:.3e}")
        assert mx < TOL_MAX and mean < TOL_MEAN, (pos, mx, mean)
    # identical inputs at different batch positions give identical outputs (no dependence on tile scheduling)
    assert np.array_equal(got[0], got[3]) and np.array_equal(got[1], got[31])


@pytest.mark.parametrize("hw,net", [((384, 384), (384, 384)), ((384, 512), (384, 384)), ((512, 768), (512, 512))])
def test_beit_large_384_and_nonsquare(cuda_device, hw, net):
    """beitl16_384 at its native size, on a 4:3 image (net 512x384: generic relative-position mode, window 24 -> 24x32)
    and — ADVICE r1 — beitl16_512 on a 3:2 image (net 768x512, nrd = 5988 > 4096)."""
    import torch
    from depthmap_b200.depthmap_generation import DptBeitEngine
    from oracle import beit_dpt, synth_weights
    from synth import synth_rgb
    name = 'beitl16_512' if net == (512, 512) else 'beitl16_384'
    sd = synth_weights.make_beit_dpt_state_dict(name, seed=4)
    eng = DptBeitEngine(sd, name, cuda_device)
    imgs = [synth_rgb(hw[0], hw[1], 90 + s) for s in range(2)]
    got = eng.forward_batch(torch.from_numpy(np.stack(imgs)).to(cuda_device), net[0], net[1]).cpu().numpy()
    want = beit_dpt.get_raw_prediction(imgs[1], sd, name, net[0], net[1])[0]
    mx, mean = _err(got[1], want)
    print(f"{name} {hw} net {net}: max {mx:.3e} mean {mean:.3e}")
    assert mx < TOL_MAX and mean < TOL_MEAN, (mx, mean)
