"""Depth-network workloads for bench.py (kept separate so bench.py stays importable before the tensor-core path exists)."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


class Dav2Stereo:
    """BASELINE.json configs[2]: depth_anything_v2 vitl @518x518 + SBS stereo (divergence 2.5, polylines fill) + normal map,
    batch 64 per GPU."""
    name = "dav2_stereo"
    encoder = "vitl"
    H = W = 518
    B = 64
    dtype = "fp16"  # tensor-core operands fp16, fp32 accumulate / residual stream (the reference's GPU path is .half())
    fill = "polylines_sharp"
    FLOP_PER_IMAGE = 1304.2e9  # SURVEY §8d, cross-checked there against FlopCounterMode on the reference module

    def __init__(self, dev, rank):
        import torch
        from bench import make_images
        from depthmap_b200.depthmap_generation import DepthAnythingV2Engine
        from oracle import synth_weights  # synthetic checkpoint-layout weights (data generation, not compute)
        self.dev = dev
        sd = synth_weights.make_dav2_state_dict(self.encoder, seed=0)
        self.engine = DepthAnythingV2Engine(sd, self.encoder, dev)
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
        n0 = self.engine.ops.launches
        if time_kernel:
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.engine.probe = {'fc1': self._ev}
        pred = self.engine.forward_batch(rgb, self.W)
        self.engine.probe = None
        depth = normalize_prediction_batch(pred, False)
        sbs = create_stereoimages_batch(rgb, depth, 2.5, 0.0, ['left-right'], 0.0, 1.0, self.fill)[0]
        normal = create_normalmap_batch(depth)
        self.launches_per_step = (self.engine.ops.launches - n0) + 3 + 3 + 1
        return depth, sbs, normal

    def step_resident(self, time_kernel=False):
        return self.step(self.rgb, time_kernel)

    def step_e2e(self):
        rgb = self.rgb_h.to(self.dev, non_blocking=True)
        outs = self.step(rgb)
        return [o.to("cpu", non_blocking=True) for o in outs]

    def e2e_bytes(self):
        return self.rgb_h.numel(), self.B * self.H * self.W * (2 + 6 + 3)

    def roofline(self, peaks, kernel_ms):
        achieved = self.fc1_flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "gemm_tcgen05_2sm_kernel (cta_group::2, 256x256 tile pair; block-0 MLP fc1: M=B*1370, N=4096, K=1024, GELU epilogue)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                "traffic": None, "peak_source": peaks["source"] + " (cuBLAS bf16 burst)", "algorithmic_flops_per_launch": self.fc1_flops,
                "kernel_ms": kernel_ms}

    def extra(self, ms_step, peaks):
        fwd = self.FLOP_PER_IMAGE * self.B / (ms_step * 1e-3) / 1e12
        return {"whole_step_tflops": fwd, "whole_step_frac_of_sustained_peak": fwd / peaks["bf16_tflops_sustained"]}

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


class DepthBeit512:
    """BASELINE.json configs[1]: dpt_beit_large_512 @512x512, batch 32 per GPU, depth only (prediction -> 16-bit depth)."""
    name = "depth_beit512"
    model = "beitl16_512"
    H = W = 512
    B = 32
    dtype = "fp16"
    FLOP_PER_IMAGE = 962.7e9

    def __init__(self, dev, rank):
        import torch
        from bench import make_images
        from depthmap_b200.depthmap_generation import DptBeitEngine
        from oracle import synth_weights  # synthetic checkpoint-layout weights (data generation, not compute)
        self.dev = dev
        sd = synth_weights.make_beit_dpt_state_dict(self.model, seed=0)
        self.engine = DptBeitEngine(sd, self.model, dev)
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
        n0 = self.engine.ops.launches
        if time_kernel:
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.engine.probe = {'fc1': self._ev}
        pred = self.engine.forward_batch(rgb, self.W, self.H)
        self.engine.probe = None
        depth = normalize_prediction_batch(pred, False)
        self.launches_per_step = (self.engine.ops.launches - n0) + 3
        return (depth,)

    def step_resident(self, time_kernel=False):
        return self.step(self.rgb, time_kernel)

    def step_e2e(self):
        rgb = self.rgb_h.to(self.dev, non_blocking=True)
        outs = self.step(rgb)
        return [o.to("cpu", non_blocking=True) for o in outs]

    def e2e_bytes(self):
        return self.rgb_h.numel(), self.B * self.H * self.W * 2

    def roofline(self, peaks, kernel_ms):
        achieved = self.fc1_flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "gemm_tcgen05_2sm_kernel (cta_group::2, 256x256 tile pair; block-0 MLP fc1: M=B*1025, N=4096, K=1024, GELU epilogue)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch (B = 32), profiles/r01_ncu_gemm_2sm_fc1.txt;
                # algorithmic: A 67 MB + W 8 MB in, C 269 MB out
                "traffic": 316.4e6 if self.B == 32 else None, "traffic_unit": "bytes/launch (ncu --set full)",
                "peak_source": peaks["source"] + " (cuBLAS bf16 burst)", "algorithmic_flops_per_launch": self.fc1_flops,
                "kernel_ms": kernel_ms}

    def extra(self, ms_step, peaks):
        fwd = self.FLOP_PER_IMAGE * self.B / (ms_step * 1e-3) / 1e12
        return {"whole_step_tflops": fwd, "whole_step_frac_of_sustained_peak": fwd / peaks["bf16_tflops_sustained"]}

    def cpu_sample(self, nthreads):
        """reference CPU path on one image: fp32 torch forward (oracle restatement of DPT-BEiT) + C oracle normalise."""
        import torch
        from oracle import beit_dpt
        from oracle import normalmap as onm
        from oracle import synth_weights
        if not hasattr(self, "_sd_cpu"):
            self._sd_cpu = synth_weights.make_beit_dpt_state_dict(self.model, seed=0)
        rgb = self.rgb_h[0].numpy()
        torch.set_num_threads(nthreads)
        t0 = time.perf_counter()
        pred, inv = beit_dpt.get_raw_prediction(rgb, self._sd_cpu, self.model, self.W, self.H)
        onm.normalize_to_u16(pred, inv)
        return 1, time.perf_counter() - t0


class Dav2StereoSmall(Dav2Stereo):
    """ViT-S variant for quick functional runs."""
    name = "dav2s_stereo"
    encoder = "vits"
    B = 8
    FLOP_PER_IMAGE = 0.0


MODEL_WORKLOADS = {"depth_beit512": DepthBeit512, "dav2_stereo": Dav2Stereo, "dav2s_stereo": Dav2StereoSmall}
