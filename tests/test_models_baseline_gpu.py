"""GPU: the depth networks on the BASELINE.json configurations (not only the small structural ones):
  C3 core  Depth-Anything-V2 ViT-L @518x518, batch > 1
  C2       dpt_beit_large_512 @512x512 at the bench batch (32) — compares images spread over the batch, because tile
           scheduling / L2 grouping of the persistent GEMMs depends on the batch
  C4 core  dpt_beit_large_384 (the ZoeDepth-NK trunk) incl. a non-square net (generic relative-position mode)
Bar: tests/precision.py (north_star's 1e-3, with the reference's own GPU precision policy measured in the same test)."""
import numpy as np
import pytest

import precision

pytestmark = pytest.mark.gpu


def test_dav2_vitl_518(cuda_device):
    import torch
    from depthmap_b200.depthmap_generation import DepthAnythingV2Engine
    from oracle import dav2 as odav2
    from oracle import synth_weights
    from synth import synth_rgb
    sd = synth_weights.make_dav2_state_dict('vitl', seed=1)
    eng = DepthAnythingV2Engine(sd, 'vitl', cuda_device)
    imgs = [synth_rgb(518, 518, 50 + s) for s in range(3)]
    got = eng.forward_batch(torch.from_numpy(np.stack(imgs)).to(cuda_device), 518).cpu().numpy()
    for i in (0, 2):
        want, _ = odav2.get_raw_prediction(imgs[i], sd, 'vitl', 518)
        assert want.max() - want.min() > 0.1
        ref16 = precision.reference_fp16_error('dav2', imgs[i], sd, 'vitl', 518, want, cuda_device)
        precision.check(f"dav2 vitl 518 B=3 img{i}", got[i], want, ref16)


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
    # a batch of one gives bit-identical results: no reduction order depends on the batch size or on the tile schedule
    solo = eng.forward_batch(torch.from_numpy(uniq[0][None]).to(cuda_device), 512, 512).cpu().numpy()[0]
    print("[precision] beitl16_512 B=32 vs B=1, image 0: max abs diff", float(np.abs(solo - got[0]).max()))
    wants, failures = {}, []
    for pos in (0, 13, 31):
        k = order[pos]
        if k not in wants:
            w = beit_dpt.get_raw_prediction(uniq[k], sd, 'beitl16_512', 512, 512)[0]
            wants[k] = (w, precision.reference_fp16_error('beit', uniq[k], sd, 'beitl16_512', (512, 512), w, cuda_device))
        try:
            precision.check(f"beitl16_512 B=32 img{pos}", got[pos], wants[k][0], wants[k][1])
        except AssertionError as e:
            failures.append(str(e))
    assert np.array_equal(solo, got[0])
    # identical inputs at different batch positions give identical outputs
    assert np.array_equal(got[0], got[3]) and np.array_equal(got[1], got[30]) and np.array_equal(got[2], got[31])
    assert not failures, failures


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
    ref16 = precision.reference_fp16_error('beit', imgs[1], sd, name, net, want, cuda_device)
    precision.check(f"{name} {hw} net {net}", got[1], want, ref16)
