"""GPU: Depth-Anything-V2 forward on the tensor-core path vs the fp32 oracle (pinned to the reference module).

Tolerance (north_star: 1e-3 on depth): the bar is on the NORMALISED depth the funnel produces, |d_gpu - d_ref| with both
maps scaled by the oracle's (max - min); fp16 operands with fp32 accumulation / fp32 residual stream."""
import numpy as np
import pytest

import precision

pytestmark = pytest.mark.gpu


def _norm_err(got, want):
    rng = float(want.max() - want.min())
    return float(np.abs(got - want).max()) / max(rng, 1e-9), float(np.abs(got - want).mean()) / max(rng, 1e-9)


@pytest.mark.parametrize("encoder,hw,net", [("vits", (70, 98), 70), ("vits", (64, 64), 56), ("vits", (120, 90), 140), ("vitb", (84, 84), 84)])
def test_dav2_forward_vs_oracle(cuda_device, encoder, hw, net):
    import torch
    from depthmap_b200.depthmap_generation import DepthAnythingV2Engine
    from oracle import dav2 as odav2
    from oracle import synth_weights
    from synth import synth_rgb
    sd = synth_weights.make_dav2_state_dict(encoder, seed=1)
    eng = DepthAnythingV2Engine(sd, encoder, cuda_device)
    imgs = [synth_rgb(hw[0], hw[1], s) for s in (3, 4)]
    got = eng.forward_batch(torch.from_numpy(np.stack(imgs)).to(cuda_device), net).cpu().numpy()
    for i, img in enumerate(imgs):
        want, inv = odav2.get_raw_prediction(img, sd, encoder, net)
        assert want.max() - want.min() > 0.1
        ref16 = precision.reference_fp16_error('dav2', img, sd, encoder, net, want, cuda_device)
        precision.check(f"dav2 {encoder} {hw} net {net} img{i}", got[i], want, ref16)


def test_modelholder_api(cuda_device):
    from PIL import Image
    from depthmap_b200.depthmap_generation import ModelHolder
    from oracle import dav2 as odav2
    from oracle import synth_weights
    from synth import synth_rgb
    sd = synth_weights.make_dav2_state_dict('vits', seed=2)
    mh = ModelHolder()
    mh.update_settings(no_half=False, precision="autocast", boost_rmax=1600)
    mh.weights_provider = lambda t: sd
    mh.ensure_models(12, cuda_device, False)
    assert mh.depth_model_type == 12 and ModelHolder.get_default_net_size(12) == [518, 518]
    img = synth_rgb(56, 84, 9)
    pred, invert = mh.get_raw_prediction(Image.fromarray(img), 56, 56)
    assert pred.dtype == np.float32 and pred.shape == (56, 84) and invert is False
    want, _ = odav2.get_raw_prediction(img, sd, 'vits', 56)
    precision.check('modelholder vits', pred, want, precision.reference_fp16_error('dav2', img, sd, 'vits', 56, want, cuda_device))
    mh.offload(); mh.reload(); mh.unload_models()
    assert mh.depth_model is None
    with pytest.raises(NotImplementedError):
        mh.ensure_models(4, cuda_device, False)      # dpt_hybrid_384: not implemented
