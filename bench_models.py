"""Depth-network workloads for bench.py (kept separate so bench.py stays importable before the tensor-core path exists).

Timed region: the model-level C-ABI handle (NativeDepthModel -> dm_depth_forward) + the HBM-side kernels.  A second, op-level
instance of the same network (the Python engine that issues the same kernels one C call at a time) exists only for the PROBE
pass after the timed region: CUDA events around every attention launch and around the block-0 fc1 GEMM, which gives the
dominant kernel's share of the step and its roofline live, in this run."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


class _NetWorkload:
    """Shared parts of the network workloads: native model for the timed region, op-level engine for the kernel probe."""
    heads = 16
    C = 1024

    def probe(self, n=20):
        """Dominant-kernel timing, live in this run: the op-level engine runs one forward (so its buffers hold this workload's
        real activations), then the fused attention kernel of the LAST block and the fc1 GEMM are launched n times back to back
        on those buffers with CUDA events around the loop (GPU-bound: no host gaps inside the interval); the forward itself is
        timed on the native model (the thing the timed region runs)."""
        import torch
        from depthmap_b200 import _lib
        eng = self.engine
        for _ in range(2):
            self._engine_forward()
        torch.cuda.synchronize()
        b = eng._bufs
        cfg = eng.cfg
        C, heads = cfg['embed_dim'], cfg['heads']
        _, nh, nw = eng._buf_key
        gh, gw = nh // eng.PATCH, nw // eng.PATCH
        N = gh * gw + 1
        last = cfg['depth'] - 1
        blk = eng.w['blocks'][last]
        PB = getattr(self, 'probe_batch', self.B)
        rows = PB * N

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        attn_ms = timed(lambda: eng.attention(last, b, PB, N, heads, C, gh, gw))
        fc1_ms = timed(lambda: eng.ops.gemm(b['h'], C, blk['fc1_w'], C, rows, 4 * C, C, act=_lib.ACT_GELU, bias=blk['fc1_b'], C=b['mlp'], ldc=4 * C))
        fwd_ms = timed(self._model_forward)
        return {"attn_ms_per_launch": float(attn_ms), "attn_launches": cfg['depth'], "fc1_ms": float(fc1_ms), "forward_ms": float(fwd_ms)}

    def roofline_from_probe(self, peaks, pr, ms_step):
        """Dominant kernel = the fused attention kernel (largest share of the step, measured here); fc1 reported beside it."""
        N = self.N
        attn_flops = 4.0 * getattr(self, 'probe_batch', self.B) * self.heads * N * N * 64          # QK^T + PV, 2 flops per MAC (2.15 + 2.15 GFLOP / image / block)
        a = attn_flops / (pr["attn_ms_per_launch"] * 1e-3) / 1e12
        f = self.fc1_flops / (pr["fc1_ms"] * 1e-3) / 1e12
        share = pr["attn_ms_per_launch"] * pr["attn_launches"] / pr["forward_ms"]
        return {"bound": "tensor", "kernel": self.ATTN_KERNEL, "achieved": a, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": a / peaks["bf16_tflops"], "traffic": None, "traffic_note": "see profiles/r02_ncu_attention_fwd4.txt (one ncu --set full capture per kernel)",
                "peak_source": peaks["source"] + " (cuBLAS bf16 burst)", "algorithmic_flops_per_launch": attn_flops,
                "kernel_ms": pr["attn_ms_per_launch"], "launches_per_step": pr["attn_launches"],
                "share_of_step": share, "share_note": "attention launches per forward x kernel time / forward time of the model handle, all CUDA events, this run",
                "secondary": {"kernel": self.FC1_KERNEL, "achieved": f, "frac": f / peaks["bf16_tflops"], "kernel_ms": pr["fc1_ms"],
                              "algorithmic_flops_per_launch": self.fc1_flops}}

    def _model_forward(self):
        return self.model.forward_batch(self.rgb, self.W, self.H)

    def extra(self, ms_step, peaks):
        fwd = self.FLOP_PER_IMAGE * self.B / (ms_step * 1e-3) / 1e12
        return {"whole_step_tflops": fwd, "whole_step_frac_of_sustained_peak": fwd / peaks["bf16_tflops_sustained"]}

    def funnel_e2e(self, steps=2):
        """The reference-facing entry point itself: PIL images in, PIL images out through core_generation_funnel (src/core.py:83)."""
        import torch
        from PIL import Image
        from depthmap_b200 import core
        holder = core.get_model_holder()
        holder.unload_models()
        holder.weights_provider = lambda t: self._state_dict()
        imgs = [Image.fromarray(self.rgb_h[i].numpy()) for i in range(self.B)]
        inp = self.funnel_options()
        n_out = 0
        times = []
        for it in range(steps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_out = sum(1 for _ in core.core_generation_funnel(None, list(imgs), None, None, inp, ops={}))
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        holder.unload_models()
        holder.weights_provider = None
        dt = float(np.mean(times[1:]))             # first pass loads / packs the model, as the reference's first call does
        return {"value": self.B / dt, "unit": "images/s", "ms_per_step": dt * 1e3, "results_per_step": n_out,
                "note": "core_generation_funnel(PIL in -> PIL out), wall clock incl. PIL<->tensor conversion, H2D/D2H and the funnel's bounded batches"}


class Dav2Stereo(_NetWorkload):
    """BASELINE.json configs[2]: depth_anything_v2 vitl @518x518 + SBS stereo (divergence 2.5, polylines fill) + normal map,
    batch 64 per GPU."""
    name = "dav2_stereo"
    encoder = "vitl"
    H = W = 518
    B = 64
    dtype = "fp16"  # tensor-core operands fp16, fp32 accumulate / residual stream (the reference's GPU path is .half())
    fill = "polylines_sharp"
    FLOP_PER_IMAGE = 1304.2e9  # SURVEY §8d, cross-checked there against FlopCounterMode on the reference module
    MODEL_TYPE = 14
    ATTN_KERNEL = "attention_fwd4_kernel<0> (fused softmax(QK^T)V, tcgen05 + TMEM, N = 1370, 16 heads x 64)"
    FC1_KERNEL = "gemm_tcgen05_2sm_kernel (cta_group::2, 256x256 tile pair; block-0 MLP fc1: M=B*1370, N=4096, K=1024, GELU epilogue)"

    def _state_dict(self):
        from oracle import synth_weights  # synthetic checkpoint-layout weights (data generation, not compute)
        return synth_weights.make_dav2_state_dict(self.encoder, seed=0)

    def funnel_options(self):
        return dict(compute_device='GPU', model_type=self.MODEL_TYPE, net_width=self.W, net_height=self.H, do_output_depth=True,
                    gen_stereo=True, stereo_modes=['left-right'], stereo_divergence=2.5, stereo_fill_algo=self.fill, gen_normalmap=True)

    def __init__(self, dev, rank):
        import torch
        from bench import make_images
        from depthmap_b200.depthmap_generation import DepthAnythingV2Engine, NativeDepthModel
        self.dev = dev
        sd = self._state_dict()
        self.model = NativeDepthModel(sd, self.MODEL_TYPE, dev)
        self.engine = DepthAnythingV2Engine(sd, self.encoder, dev)      # probe pass only
        del sd
        rgb, _ = make_images(self.B, self.H, self.W, rank)
        self.rgb_h = torch.from_numpy(rgb).pin_memory()
        self.rgb = self.rgb_h.to(dev)
        cfg = self.engine.cfg
        C = cfg['embed_dim']
        self.N = (self.H // 14) * (self.W // 14) + 1
        self.fc1_flops = 2.0 * self.B * self.N * C * 4 * C
        self.launches_per_step = None

    def config(self):
        return {"workload": "depth_anything_v2 vitl 518x518 -> u16 depth -> SBS stereo (divergence 2.5, polylines_sharp) -> "
                            "normal map (Sobel 3)", "batch_per_gpu": self.B, "height": self.H, "width": self.W,
                "weights": "seeded synthetic, upstream checkpoint layout",
                "l2_policy": "activations per step (>10 GB) far exceed the 126 MB L2"}

    def step(self, rgb, time_kernel=False):
        import torch
        from depthmap_b200.core import normalize_prediction_batch
        from depthmap_b200.normalmap_generation import create_normalmap_batch
        from depthmap_b200.stereoimage_generation import create_stereoimages_batch
        n0 = self.model.launches
        pred = self.model.forward_batch(rgb, self.W, self.H)
        depth = normalize_prediction_batch(pred, False)
        sbs = create_stereoimages_batch(rgb, depth, 2.5, 0.0, ['left-right'], 0.0, 1.0, self.fill)[0]
        normal = create_normalmap_batch(depth)
        n1 = self.model.launches
        if n1 > n0:
            self.launches_per_step = (n1 - n0) + 3 + 3 + 1
        return depth, sbs, normal

    def _engine_forward(self):
        return self.engine.forward_batch(self.rgb, self.W)

    def step_resident(self, time_kernel=False):
        return self.step(self.rgb, time_kernel)

    def step_e2e(self):
        rgb = self.rgb_h.to(self.dev, non_blocking=True)
        outs = self.step(rgb)
        return [o.to("cpu", non_blocking=True) for o in outs]

    def e2e_bytes(self):
        return self.rgb_h.numel(), self.B * self.H * self.W * (2 + 6 + 3)

    def cpu_sample(self, nthreads):
        """reference CPU path on one image: fp32 torch forward (oracle restatement) + C oracle stereo / normal map."""
        import torch
        from oracle import dav2 as odav2
        from oracle import normalmap as onm
        from oracle import stereo as ost
        from oracle import synth_weights
        if not hasattr(self, "_sd_cpu"):
            self._sd_cpu = synth_weights.make_dav2_state_dict(self.encoder, seed=0)
        rgb = self.rgb_h[0].numpy()
        torch.set_num_threads(nthreads)
        t0 = time.perf_counter()
        pred, inv = odav2.get_raw_prediction(rgb, self._sd_cpu, self.encoder, self.W)
        d = onm.normalize_to_u16(pred, inv)
        ost.create_stereoimages(rgb, d, 2.5, 0.0, ['left-right'], 0.0, 1.0, self.fill, return_arrays=True, nthreads=nthreads)
        onm.create_normalmap(d, return_array=True)
        return 1, time.perf_counter() - t0


class DepthBeit512(_NetWorkload):
    """BASELINE.json configs[1]: dpt_beit_large_512 @512x512, batch 32 per GPU, depth only (prediction -> 16-bit depth)."""
    name = "depth_beit512"
    model_name = "beitl16_512"
    H = W = 512
    B = 32
    dtype = "fp16"
    FLOP_PER_IMAGE = 962.7e9
    MODEL_TYPE = 1
    ATTN_KERNEL = "attention_fwd4_kernel<3> (fused softmax(QK^T + rel-pos bias)V, tcgen05 + TMEM, N = 1025, 16 heads x 64; + class-row kernel)"
    FC1_KERNEL = "gemm_tcgen05_2sm_kernel (cta_group::2, 256x256 tile pair; block-0 MLP fc1: M=B*1025, N=4096, K=1024, GELU epilogue)"

    def _state_dict(self):
        from oracle import synth_weights  # synthetic checkpoint-layout weights (data generation, not compute)
        return synth_weights.make_beit_dpt_state_dict(self.model_name, seed=0)

    def funnel_options(self):
        return dict(compute_device='GPU', model_type=self.MODEL_TYPE, net_width=self.W, net_height=self.H, do_output_depth=True)

    def __init__(self, dev, rank):
        import torch
        from bench import make_images
        from depthmap_b200.depthmap_generation import DptBeitEngine, NativeDepthModel
        self.dev = dev
        sd = self._state_dict()
        self.model = NativeDepthModel(sd, self.MODEL_TYPE, dev)
        self.engine = DptBeitEngine(sd, self.model_name, dev)           # probe pass only
        del sd
        rgb, _ = make_images(self.B, self.H, self.W, rank)
        self.rgb_h = torch.from_numpy(rgb).pin_memory()
        self.rgb = self.rgb_h.to(dev)
        C = self.engine.cfg['embed_dim']
        self.N = (self.H // 16) * (self.W // 16) + 1
        self.fc1_flops = 2.0 * self.B * self.N * C * 4 * C
        self.launches_per_step = None

    def config(self):
        return {"workload": "dpt_beit_large_512 (MiDaS 3.1) 512x512 -> float32 prediction -> u16 depth", "batch_per_gpu": self.B,
                "height": self.H, "width": self.W, "weights": "seeded synthetic, MiDaS checkpoint layout",
                "l2_policy": "activations per step (>5 GB) far exceed the 126 MB L2"}

    def step(self, rgb, time_kernel=False):
        import torch
        from depthmap_b200.core import normalize_prediction_batch
        n0 = self.model.launches
        pred = self.model.forward_batch(rgb, self.W, self.H)
        depth = normalize_prediction_batch(pred, False)
        n1 = self.model.launches
        if n1 > n0:
            self.launches_per_step = (n1 - n0) + 3
        return (depth,)

    def _engine_forward(self):
        return self.engine.forward_batch(self.rgb, self.W, self.H)

    def step_resident(self, time_kernel=False):
        return self.step(self.rgb, time_kernel)

    def step_e2e(self):
        rgb = self.rgb_h.to(self.dev, non_blocking=True)
        outs = self.step(rgb)
        return [o.to("cpu", non_blocking=True) for o in outs]

    def e2e_bytes(self):
        return self.rgb_h.numel(), self.B * self.H * self.W * 2

    def cpu_sample(self, nthreads):
        """reference CPU path on one image: fp32 torch forward (oracle restatement of DPT-BEiT) + C oracle normalise."""
        import torch
        from oracle import beit_dpt
        from oracle import normalmap as onm
        from oracle import synth_weights
        if not hasattr(self, "_sd_cpu"):
            self._sd_cpu = synth_weights.make_beit_dpt_state_dict(self.model_name, seed=0)
        rgb = self.rgb_h[0].numpy()
        torch.set_num_threads(nthreads)
        t0 = time.perf_counter()
        pred, inv = beit_dpt.get_raw_prediction(rgb, self._sd_cpu, self.model_name, self.W, self.H)
        onm.normalize_to_u16(pred, inv)
        return 1, time.perf_counter() - t0


class ZoeAnaglyph(_NetWorkload):
    """BASELINE.json configs[3]: zoedepth_nk at 768x768 + anaglyph stereo, 32 images per GPU (256 over 8 GPUs, sharded, one
    NCCL gather).  Every image runs twice through the DPT-BEiT-L-384 core (flip TTA): 64 forwards at the 512x512 net size
    that PrepForMidas picks for the reflect-padded 884x884 input with the UI default net (w 384, h 512)."""
    name = "zoedepth_nk768"
    H = W = 768
    B = 32
    dtype = "fp16"
    NET_W, NET_H = 384, 512
    FLOP_PER_IMAGE = 2 * 962.7e9 + 2 * 10e9      # two core forwards at 512x512 + the metric head (~1%)
    ATTN_KERNEL = "attention_fwd4_kernel<3> (fused softmax(QK^T + rel-pos bias)V of the BEiT-L-384 core at a 32x32 window, 64 forwards)"
    FC1_KERNEL = "gemm_tcgen05_2sm_kernel (block MLP fc1: M=64*1025, N=4096, K=1024, GELU epilogue)"

    def _state_dict(self):
        from oracle import beit_dpt, synth_weights
        csd = synth_weights.make_beit_dpt_state_dict('beitl16_384', seed=0)
        hsd = synth_weights.make_zoedepth_head_state_dict(feat_ch=beit_dpt.CONFIGS['beitl16_384']['features'], seed=100, gain=1.5)
        sd = {"core.core." + k: v for k, v in csd.items()}
        sd.update(hsd)
        return sd

    def funnel_options(self):
        return dict(compute_device='GPU', model_type=9, net_width=self.NET_W, net_height=self.NET_H, do_output_depth=True, gen_stereo=True,
                    stereo_modes=['red-cyan-anaglyph'], stereo_divergence=2.5, stereo_fill_algo='polylines_sharp')

    def __init__(self, dev, rank):
        import torch
        from bench import make_images
        from depthmap_b200.depthmap_generation import ZoeDepthNKEngine
        self.dev = dev
        self.engine = ZoeDepthNKEngine(self._state_dict(), dev)
        self.model = self.engine
        rgb, _ = make_images(self.B, self.H, self.W, rank)
        self.rgb_h = torch.from_numpy(rgb).pin_memory()
        self.rgb = self.rgb_h.to(dev)
        self.probe_batch = 2 * self.B
        self.N = 32 * 32 + 1
        self.fc1_flops = 2.0 * self.probe_batch * self.N * 1024 * 4096
        self.launches_per_step = None

    def config(self):
        return {"workload": "zoedepth_nk (DPT-BEiT-L-384 core + metric head, pad + flip TTA) 768x768 -> u16 depth -> red-cyan anaglyph "
                            "(divergence 2.5, polylines_sharp)", "batch_per_gpu": self.B, "height": self.H, "width": self.W,
                "net": "384x512 (UI default) -> 512x512 for the padded 884x884 input", "weights": "seeded synthetic, ZoeD_M12_NK checkpoint layout",
                "l2_policy": "activations per step (>20 GB) far exceed the 126 MB L2"}

    def step(self, rgb, time_kernel=False):
        from depthmap_b200.core import normalize_prediction_batch
        from depthmap_b200.stereoimage_generation import create_stereoimages_batch
        n0 = self.engine.ops.launches
        pred = self.engine.forward_batch(rgb, self.NET_W, self.NET_H)
        depth = normalize_prediction_batch(pred, True)
        ana = create_stereoimages_batch(rgb, depth, 2.5, 0.0, ['red-cyan-anaglyph'], 0.0, 1.0, 'polylines_sharp')[0]
        self.launches_per_step = (self.engine.ops.launches - n0) + 3 + 3
        return depth, ana

    def step_resident(self, time_kernel=False):
        return self.step(self.rgb, time_kernel)

    def step_e2e(self):
        rgb = self.rgb_h.to(self.dev, non_blocking=True)
        outs = self.step(rgb)
        return [o.to("cpu", non_blocking=True) for o in outs]

    def e2e_bytes(self):
        return self.rgb_h.numel(), self.B * self.H * self.W * (2 + 3)

    def _engine_forward(self):
        return self.engine.forward_batch(self.rgb, self.NET_W, self.NET_H)

    def _model_forward(self):
        return self.engine.forward_batch(self.rgb, self.NET_W, self.NET_H)

    def cpu_sample(self, nthreads):
        """reference CPU path on one image: fp32 torch ZoeDepth-NK (oracle restatement, 2 core forwards) + C oracle normalise / stereo."""
        import torch
        from oracle import normalmap as onm
        from oracle import stereo as ost
        from oracle import zoedepth as ozd
        if not hasattr(self, "_sd_cpu"):
            self._sd_cpu = self._state_dict()
        rgb = self.rgb_h[0].numpy()
        torch.set_num_threads(nthreads)
        t0 = time.perf_counter()
        pred, inv = ozd.get_raw_prediction(rgb, self._sd_cpu, self.NET_W, self.NET_H, core_name='beitl16_384')
        d = onm.normalize_to_u16(pred, inv)
        ost.create_stereoimages(rgb, d, 2.5, 0.0, ['red-cyan-anaglyph'], 0.0, 1.0, 'polylines_sharp', return_arrays=True, nthreads=nthreads)
        return 1, time.perf_counter() - t0


class Dav2StereoSmall(Dav2Stereo):
    """ViT-S variant for quick functional runs."""
    name = "dav2s_stereo"
    encoder = "vits"
    B = 8
    FLOP_PER_IMAGE = 0.0
    MODEL_TYPE = 12
    heads, C = 6, 384


MODEL_WORKLOADS = {"depth_beit512": DepthBeit512, "dav2_stereo": Dav2Stereo, "dav2s_stereo": Dav2StereoSmall, "zoedepth_nk768": ZoeAnaglyph}
