"""GPU: the drop-in entry point core_generation_funnel / run_depthmap (reference src/core.py:83-336) driven end to end
with PIL inputs: kinds, yield order (src/core.py:194-305), payload types, and the payloads against the oracle:
  * u16 depth: the oracle's normalisation of OUR float32 prediction must reproduce our u16 bit for bit (N1), and the
    prediction itself is within the network tolerance of the fp32 oracle network;
  * stereo / normal map: bit-exact against the oracle applied to OUR u16 depth (S1-S5, M1);
  * custom depth map branch, degenerate (flat) image, mixed sizes, bounded batches, OOM re-raise."""
import numpy as np
import pytest
from PIL import Image

from synth import synth_depth_u16, synth_rgb

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture()
def funnel(cuda_device):
    from depthmap_b200 import core
    from oracle import synth_weights
    sd = synth_weights.make_dav2_state_dict('vits', seed=2)
    holder = core.get_model_holder()
    holder.unload_models()
    holder.weights_provider = lambda t: sd
    yield core, sd
    holder.unload_models()
    holder.weights_provider = None


def _opts(**kw):
    d = dict(compute_device='GPU', model_type=12, net_width=70, net_height=70, net_size_match=False, boost=False,
             do_output_depth=True, gen_stereo=False, gen_normalmap=False)
    d.update(kw)
    return d


def test_funnel_full_path_matches_oracle(funnel):
    core, sd = funnel
    from oracle import dav2 as odav2
    from oracle import normalmap as onm
    from oracle import stereo as ost
    sizes = [(70, 98), (70, 98), (84, 70), (70, 98)]          # two runs of equal size around an odd one
    rgbs = [synth_rgb(h, w, 20 + i) for i, (h, w) in enumerate(sizes)]
    imgs = [Image.fromarray(x) for x in rgbs]
    modes = ['left-right', 'red-cyan-anaglyph']
    inp = _opts(gen_stereo=True, stereo_modes=modes, gen_normalmap=True, do_output_depth_prediction=True,
                stereo_divergence=3.0, stereo_fill_algo='polylines_sharp')
    out = list(core.core_generation_funnel(None, imgs, None, None, inp, ops={}))
    # order and kinds: per image depth_prediction, depth, one per stereo mode, normalmap; images ascending
    want_kinds = ['depth_prediction', 'depth'] + modes + ['normalmap']
    assert [(i, k) for i, k, _ in out] == [(i, k) for i in range(len(imgs)) for k in want_kinds]
    for i in range(len(imgs)):
        res = {k: v for j, k, v in out if j == i}
        pred = res['depth_prediction']
        assert isinstance(pred, np.ndarray) and pred.dtype == np.float32 and pred.shape == sizes[i]
        want_pred, inv = odav2.get_raw_prediction(rgbs[i], sd, 'vits', 70)
        rng = float(want_pred.max() - want_pred.min())
        assert float(np.abs(pred - want_pred).max()) / rng < 2 * TOL
        depth = res['depth']
        assert isinstance(depth, Image.Image) and depth.mode in ('I;16', 'I;16L', 'I;16N')
        d16 = np.asarray(depth).astype(np.uint16)
        assert np.array_equal(d16, onm.normalize_to_u16(pred, False))              # N1, bit-exact given the prediction
        want_st = ost.create_stereoimages(rgbs[i], d16, 3.0, 0.0, modes, 0.0, 1.0, 'polylines_sharp')
        for m, w in zip(modes, want_st):
            assert isinstance(res[m], Image.Image) and res[m].mode == 'RGB'
            assert np.array_equal(np.asarray(res[m]), np.asarray(w)), (i, m)
        assert np.array_equal(np.asarray(res['normalmap']), onm.create_normalmap(d16, return_array=True))


def test_funnel_is_lazy_and_bounded(funnel, monkeypatch):
    core, sd = funnel
    monkeypatch.setenv("DEPTHMAP_B200_MAX_BATCH", "2")
    imgs = [Image.fromarray(synth_rgb(56, 56, 40 + i)) for i in range(5)]
    batches = []
    holder = core.get_model_holder()
    orig = type(holder).get_raw_prediction_batch

    def spy(self, rgb, nw, nh):
        batches.append(int(rgb.shape[0]))
        return orig(self, rgb, nw, nh)
    monkeypatch.setattr(type(holder), "get_raw_prediction_batch", spy)
    gen = core.run_depthmap(None, imgs, None, None, _opts(net_width=56, net_height=56), ops={})
    first = next(gen)
    assert first[0] == 0 and first[1] == 'depth' and batches == [2]      # only the first batch has run
    rest = list(gen)
    assert batches == [2, 2, 1] and [i for i, _, _ in rest] == [1, 2, 3, 4]


def test_funnel_custom_depthmap_and_options(funnel):
    core, sd = funnel
    from oracle import normalmap as onm
    from oracle import stereo as ost
    rgb = synth_rgb(40, 64, 7)
    d16 = synth_depth_u16(40, 64, 7)
    # 16-bit single-channel PIL depth map: out = asarray(float) / 2**16 -> convert_to_i16 (src/core.py:146-174,44-50)
    dm = Image.fromarray(d16)
    inp = _opts(gen_stereo=True, stereo_modes=['top-bottom'], stereo_fill_algo='naive_interpolating', gen_normalmap=True,
                normalmap_invert=True, output_depth_invert=True, output_depth_combine=True,
                output_depth_combine_axis='Vertical', do_output_depth_prediction=True)
    out = list(core.core_generation_funnel(None, [Image.fromarray(rgb)], [dm], None, inp, ops={}))
    assert [k for _, k, _ in out] == ['concat_depth', 'top-bottom', 'normalmap']   # no depth_prediction for custom maps
    unit = np.asarray(dm, dtype="float") / 2.0 ** 16
    want16 = np.clip(unit * 65536 + 0.0001, 0, 65535.9).astype("uint16")
    res = {k: v for _, k, v in out}
    inv = np.bitwise_not(want16)
    rgbd = np.zeros_like(rgb)
    for k in range(3):
        rgbd[:, :, k] = inv / 256.0
    assert np.array_equal(np.asarray(res['concat_depth']), np.concatenate((rgb, rgbd), axis=0))
    want_st = ost.create_stereoimages(rgb, want16, 2.5, 0.0, ['top-bottom'], 0.0, 1.0, 'naive_interpolating')[0]
    assert np.array_equal(np.asarray(res['top-bottom']), np.asarray(want_st))
    assert np.array_equal(np.asarray(res['normalmap']), onm.create_normalmap(want16, None, 3, None, True, return_array=True))
    # 8-bit RGB depth map resized to the image (LANCZOS), first channel / 256
    dm8 = Image.fromarray(np.repeat((d16 >> 8).astype(np.uint8)[:, :, None], 3, axis=2)).resize((32, 20))
    out = list(core.core_generation_funnel(None, [Image.fromarray(rgb)], [dm8], None, _opts(), ops={}))
    ref = dm8.resize((64, 40), Image.Resampling.LANCZOS)
    unit = np.asarray(ref, dtype="float")[:, :, 0] / 256.0
    assert np.array_equal(np.asarray(out[0][2]).astype(np.uint16), np.clip(unit * 65536 + 0.0001, 0, 65535.9).astype("uint16"))


def test_funnel_degenerate_prediction_and_mode_I(funnel, monkeypatch):
    core, sd = funnel
    import torch
    holder = core.get_model_holder()

    def flat(self, rgb, nw, nh):
        return torch.full((rgb.shape[0], rgb.shape[1], rgb.shape[2]), 3.25, dtype=torch.float32, device=rgb.device), False
    holder.ensure_models(12, torch.device("cuda"), False)
    monkeypatch.setattr(type(holder), "get_raw_prediction_batch", flat)
    img_i = Image.fromarray((synth_depth_u16(32, 48, 3).astype(np.int32)), mode='I')
    imgs = [img_i]
    out = list(core.core_generation_funnel(None, imgs, None, None, _opts(do_output_depth_prediction=True, gen_normalmap=True), ops={}))
    assert imgs[0].mode == 'RGB'                                     # converted in the caller's list (src/core.py:135-137)
    assert [k for _, k, _ in out] == ['depth', 'normalmap']          # degenerate map: no depth_prediction, black depth
    assert not np.asarray(out[0][2]).any()
    nm = np.asarray(out[1][2])
    assert (nm == np.array([128, 128, 255], np.uint8)).all()


def test_funnel_errors(funnel, monkeypatch):
    core, sd = funnel
    holder = core.get_model_holder()
    assert list(core.core_generation_funnel(None, [], None, None, _opts())) == []
    with pytest.raises(NotImplementedError):
        list(core.core_generation_funnel(None, [Image.fromarray(synth_rgb(32, 32, 1))], None, None, _opts(gen_simple_mesh=True)))

    def oom(self, rgb, nw, nh):
        raise RuntimeError("CUDA out of memory. Tried to allocate 1.00 GiB")
    monkeypatch.setattr(type(holder), "get_raw_prediction_batch", oom)
    with pytest.raises(Exception) as ei:
        list(core.core_generation_funnel(None, [Image.fromarray(synth_rgb(32, 32, 1))], None, None, _opts(), ops={}))
    assert "out of GPU memory, could not generate depthmap" in str(ei.value)
    assert holder.offloaded                                         # `finally` offloads (src/core.py:330-334)
