"""GPU: video mode's cross-frame normalisation against the oracle restatement of
src/video_mode.py:103-128 (oracle/video.py, pinned bit-for-bit to the reference function): bit-exact, dtypes included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(n, h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.normal(size=(h, w)).astype(np.float32)
    return [(base * (1 + 0.1 * i) + rng.normal(scale=0.3, size=(h, w)).astype(np.float32) + np.float32(0.05 * i)).astype(np.float32) for i in range(n)]


@pytest.mark.parametrize("n,hw", [(1, (8, 12)), (2, (16, 16)), (7, (33, 41)), (24, (64, 48))])
@pytest.mark.parametrize("smoothening", ["none", "experimental", "other"])
def test_video_normalisation_bit_exact(cuda_device, n, hw, smoothening):
    from depthmap_b200.video_mode import process_predicitons
    from oracle import video as ov
    fr = _frames(n, hw[0], hw[1], 100 * n + hw[0])
    got = process_predicitons([f.copy() for f in fr], smoothening)
    want = ov.process_predictions([f.copy() for f in fr], smoothening)
    assert len(got) == len(want) == n
    for g, w in zip(got, want):
        assert g.dtype == w.dtype and g.shape == w.shape
        assert np.array_equal(g, w), (smoothening, n, float(np.abs(g.astype(np.float64) - w).max()))


def test_video_sharded_halo_equals_whole(cuda_device):
    """The sharded form (frames [first, first+N) of a longer clip, halo supplied) reproduces the blend of the whole clip:
    the host-side halo arithmetic, without a process group."""
    import torch
    from depthmap_b200 import _lib as L
    from depthmap_b200.video_mode import halo_plan
    lib = L.load()
    fr = np.stack(_frames(9, 20, 24, 5))
    whole = torch.from_numpy(fr).to(cuda_device)
    ref = torch.empty_like(whole)
    L.check(lib.dm_video_blend(whole.data_ptr(), 20 * 24, 0, 9, 9, 0, 9, ref.data_ptr(), L.stream_ptr()))
    for first, n in [(0, 3), (3, 1), (4, 4), (8, 1)]:
        before, after = halo_plan(first, n, 9)
        lo = before[0] if before else first
        hi = after[-1] if after else first + n - 1
        local = whole[lo:hi + 1].contiguous()
        out = torch.empty(n, 20, 24, dtype=torch.float32, device=cuda_device)
        L.check(lib.dm_video_blend(local.data_ptr(), 20 * 24, lo, local.shape[0], 9, first, n, out.data_ptr(), L.stream_ptr()))
        assert torch.equal(out, ref[first:first + n])
