"""GPU: ZoeDepth-NK (model type 9) — PIL-equivalent uint8 image in, metric depth out — against the
fp32 oracle (oracle/zoedepth.py: head + pad/flip TTA pinned to the reference module, core = oracle/beit_dpt.py).
Covers both routed heads (nyu / kitti weights chosen by the per-forward router), non-square images, the pad + flip
augmentation and the resize back; bar = tests/precision.py."""
import numpy as np
import pytest

import precision

pytestmark = pytest.mark.gpu


def make_zoe_state_dict(core, seed, gain=1.5):
    from oracle import beit_dpt, synth_weights
    csd = synth_weights.make_beit_dpt_state_dict(core, seed=seed)
    hsd = synth_weights.make_zoedepth_head_state_dict(feat_ch=beit_dpt.CONFIGS[core]['features'], seed=seed + 100, gain=gain)
    sd = {"core.core." + k: v for k, v in csd.items()}
    sd.update(hsd)
    return sd


# seeds 1 / 3 route to nyu / kitti with a logit margin > 0.1 on these inputs (probed with the oracle)
@pytest.mark.parametrize("seed,hw,net", [(1, (96, 128), (64, 64)), (3, (96, 128), (64, 64)), (3, (70, 50), (64, 96)), (1, (64, 64), (96, 96))])
def test_zoedepth_tiny_core_vs_oracle(cuda_device, seed, hw, net):
    import torch
    from depthmap_b200.depthmap_generation import ZoeDepthNKEngine
    from oracle import zoedepth as ozd
    from synth import synth_rgb
    sd = make_zoe_state_dict('beit_tiny', seed)
    eng = ZoeDepthNKEngine(sd, cuda_device, core_name='beit_tiny')
    imgs = [synth_rgb(hw[0], hw[1], 5 + s) for s in range(2)]
    got = eng.forward_batch(torch.from_numpy(np.stack(imgs)).to(cuda_device), net[0], net[1]).cpu().numpy()
    assert got.shape == (2, hw[0], hw[1]) and got.dtype == np.float32
    for i, img in enumerate(imgs):
        want, invert = ozd.get_raw_prediction(img, sd, net[0], net[1], core_name='beit_tiny')
        assert invert is True and want.max() - want.min() > 0.05
        ref16 = precision.reference_fp16_error_zoe(img, sd, net[0], net[1], 'beit_tiny', want, cuda_device)
        precision.check(f"zoedepth_nk tiny seed{seed} {hw} net {net} img{i}", got[i], want, ref16, slack=1.5)


def test_zoedepth_nk_beit384_core(cuda_device):
    """The real configuration: DPT-BEiT-L-384 core, UI default net size for model type 9 (w 384, h 512, src/depthmap_generation.py:333)."""
    import torch
    from depthmap_b200.depthmap_generation import ModelHolder
    from oracle import zoedepth as ozd
    from synth import synth_rgb
    sd = make_zoe_state_dict('beitl16_384', 3)
    mh = ModelHolder()
    mh.weights_provider = lambda t: sd
    mh.ensure_models(9, cuda_device, False)
    assert ModelHolder.get_default_net_size(9) == [384, 512]
    img = synth_rgb(384, 320, 21)
    from PIL import Image
    pred, invert = mh.get_raw_prediction(Image.fromarray(img), 384, 512)
    assert invert is True and pred.shape == (384, 320) and pred.dtype == np.float32
    want, _ = ozd.get_raw_prediction(img, sd, 384, 512, core_name='beitl16_384')
    ref16 = precision.reference_fp16_error_zoe(img, sd, 384, 512, 'beitl16_384', want, cuda_device)
    precision.check("zoedepth_nk beitl16_384 (384x320 image)", pred, want, ref16)
    mh.unload_models()
