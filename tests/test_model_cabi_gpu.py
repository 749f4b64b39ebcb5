"""GPU: the model-level C-ABI (dm_model_create / dm_depth_forward / dm_model_destroy, csrc/model.cu) against the op-level
path (the Python engines that string the same kernels together, themselves checked against the oracle): the two must agree
BIT FOR BIT — same packed weights, same launch sequence — eagerly, through the handle's own CUDA graph (3rd call) and when
recorded into an outer graph; plus shape changes, resolution-table rebuilds, error paths and B = 1 latency."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _imgs(B, h, w, seed=0):
    from synth import synth_rgb
    return np.stack([synth_rgb(h, w, seed + s) for s in range(B)])


@pytest.mark.parametrize("mtype,encoder,hw,net", [(12, 'vits', (70, 98), (70, 70)), (12, 'vits', (64, 64), (56, 56)), (13, 'vitb', (84, 84), (84, 84)),
                                                   (12, 'vits', (120, 90), (140, 140))])
def test_native_dav2_equals_op_level_path(cuda_device, mtype, encoder, hw, net):
    import torch
    from depthmap_b200.depthmap_generation import DepthAnythingV2Engine, NativeDepthModel
    from oracle import synth_weights
    sd = synth_weights.make_dav2_state_dict(encoder, seed=1)
    eng = DepthAnythingV2Engine(sd, encoder, cuda_device)
    nat = NativeDepthModel(sd, mtype, cuda_device)
    rgb = torch.from_numpy(_imgs(2, *hw)).to(cuda_device)
    want = eng.forward_batch(rgb, net[0]).cpu().numpy()
    assert nat.net_size(hw[1], hw[0], net[0], net[1]) == tuple(eng.net_size(hw[1], hw[0], net[0], net[1]))
    for call in range(4):                              # 1: eager, 2: capture + launch, 3-4: replay
        got = nat.forward_batch(rgb, net[0], net[1]).cpu().numpy()
        assert np.array_equal(got, want), (call, float(np.abs(got - want).max()))     # the position-embedding table is shared (dm_dinov2_pos_embed)
    nat.close()


@pytest.mark.parametrize("hw,net", [((64, 96), (64, 64)), ((96, 96), (96, 96)), ((80, 50), (64, 64))])
def test_native_beit_equals_op_level_path(cuda_device, hw, net):
    import torch
    from depthmap_b200.depthmap_generation import DptBeitEngine, NativeDepthModel
    from oracle import synth_weights
    sd = synth_weights.make_beit_dpt_state_dict('beit_tiny', seed=3)
    eng = DptBeitEngine(sd, 'beit_tiny', cuda_device)
    nat = NativeDepthModel(sd, -100, cuda_device)      # -100: the structural test configuration of the native model table
    rgb = torch.from_numpy(_imgs(3, *hw, seed=7)).to(cuda_device)
    want = eng.forward_batch(rgb, net[0], net[1]).cpu().numpy()
    for call in range(3):
        got = nat.forward_batch(rgb, net[0], net[1]).cpu().numpy()
        assert np.array_equal(got, want), (call, float(np.abs(got - want).max()))     # the resized bias tables are shared (dm_beit_rel_table)
    # a different batch size and resolution on the same handle, then back
    rgb2 = torch.from_numpy(_imgs(1, 96, 64, seed=9)).to(cuda_device)
    w2 = eng.forward_batch(rgb2, 64, 96).cpu().numpy()
    g2 = nat.forward_batch(rgb2, 64, 96).cpu().numpy()
    assert np.array_equal(g2, w2)
    got = nat.forward_batch(rgb, net[0], net[1]).cpu().numpy()
    assert np.array_equal(got, want)
    nat.close()


@pytest.mark.parametrize("hw,net", [((64, 64), (64, 64)), ((96, 128), (96, 96))])
def test_native_vit_equals_op_level_path(cuda_device, hw, net):
    """dpt_large_384 family (model type 3; -101 = its structural test configuration)"""
    import torch
    from depthmap_b200.depthmap_generation import DptVitEngine, NativeDepthModel
    from oracle import synth_weights
    sd = synth_weights.make_beit_dpt_state_dict('vit_tiny', seed=5)
    eng = DptVitEngine(sd, 'vit_tiny', cuda_device)
    nat = NativeDepthModel(sd, -101, cuda_device)
    rgb = torch.from_numpy(_imgs(2, *hw, seed=11)).to(cuda_device)
    want = eng.forward_batch(rgb, net[0], net[1]).cpu().numpy()
    for call in range(3):
        got = nat.forward_batch(rgb, net[0], net[1]).cpu().numpy()
        assert np.array_equal(got, want), (call, float(np.abs(got - want).max()))
    nat.close()


def test_native_beit512_bit_exact_and_outer_graph(cuda_device):
    """dpt_beit_large_512 at its native window (no table resize): the native model reproduces the op-level path exactly, also
    when its launches are recorded into a caller's CUDA graph."""
    import torch
    from depthmap_b200.depthmap_generation import DptBeitEngine, NativeDepthModel
    from oracle import synth_weights
    sd = synth_weights.make_beit_dpt_state_dict('beitl16_512', seed=3)
    eng = DptBeitEngine(sd, 'beitl16_512', cuda_device)
    nat = NativeDepthModel(sd, 1, cuda_device)
    rgb = torch.from_numpy(_imgs(2, 512, 512, seed=70)).to(cuda_device)
    want = eng.forward_batch(rgb, 512, 512)
    for _ in range(3):
        assert torch.equal(nat.forward_batch(rgb, 512, 512), want)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        nat.forward_batch(rgb, 512, 512)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = nat.forward_batch(rgb, 512, 512)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    # B = 1 latency through the reference's call shape (ModelHolder.get_raw_prediction, src/core.py:185): op-level vs handle
    one = rgb[:1].contiguous()

    def lat(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    t_ops = lat(lambda: eng.forward_batch(one, 512, 512))
    t_nat = lat(lambda: nat.forward_batch(one, 512, 512))
    print(f"[latency] dpt_beit_large_512 B=1: op-level path {t_ops:.2f} ms/forward, model handle {t_nat:.2f} ms/forward")
    assert t_nat < t_ops
    nat.close()


def test_native_model_errors(cuda_device):
    import ctypes
    import torch
    from depthmap_b200 import _lib
    from depthmap_b200.depthmap_generation import NativeDepthModel
    from oracle import synth_weights
    sd = synth_weights.make_dav2_state_dict('vits', seed=1)
    with pytest.raises(NotImplementedError):
        NativeDepthModel(sd, 7, cuda_device)
    bad = dict(sd)
    del bad['pretrained.blocks.3.attn.qkv.weight']
    with pytest.raises(ValueError) as ei:
        NativeDepthModel(bad, 12, cuda_device)
    assert 'pretrained.blocks.3.attn.qkv.weight' in str(ei.value)
    nat = NativeDepthModel(sd, 12, cuda_device)
    with pytest.raises(ValueError):
        nat.forward_batch(torch.zeros(1, 8, 8, 3, dtype=torch.float32, device=cuda_device), 56)
    nat.close()
    with pytest.raises(RuntimeError):
        nat.forward_batch(torch.zeros(1, 56, 56, 3, dtype=torch.uint8, device=cuda_device), 56)
    assert _lib.load().dm_model_destroy(None) == 0
