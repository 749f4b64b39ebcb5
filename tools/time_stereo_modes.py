"""Times the stereo kernel per fill mode at 2048x2048 (B=16) with CUDA events; prints JSON lines."""
import json, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from bench import make_images, _peaks
from depthmap_b200.core import normalize_prediction_batch
from depthmap_b200.normalmap_generation import create_normalmap_batch
from depthmap_b200.stereoimage_generation import create_stereoimages_batch
B, H, W = 16, 2048, 2048
rgb, pred = make_images(B, H, W, 0)
dev = torch.device('cuda')
rgb_t, pred_t = torch.from_numpy(rgb).to(dev), torch.from_numpy(pred).to(dev)
depth = normalize_prediction_batch(pred_t, False)
peaks = _peaks()
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for fill in ['none', 'naive', 'naive_interpolating', 'polylines_soft', 'polylines_sharp']:
    for mode, bpp in (('left-right', 11), ('red-cyan-anaglyph', 8)):
        ms = timeit(lambda: create_stereoimages_batch(rgb_t, depth, 2.5, 0.0, [mode], 0.0, 1.0, fill))
        gbs = bpp * H * W * B / ms / 1e6
        print(json.dumps({"op": "stereo", "fill": fill, "mode": mode, "ms": ms, "GBps": gbs, "frac_hbm": gbs / peaks['hbm_gbs']}))
ms = timeit(lambda: create_normalmap_batch(depth))
print(json.dumps({"op": "normalmap_sobel3", "ms": ms, "GBps": 5 * H * W * B / ms / 1e6, "frac_hbm": 5 * H * W * B / ms / 1e6 / peaks['hbm_gbs']}))
ms = timeit(lambda: normalize_prediction_batch(pred_t, False))
print(json.dumps({"op": "normalize_u16", "ms": ms, "GBps": 10 * H * W * B / ms / 1e6, "frac_hbm": 10 * H * W * B / ms / 1e6 / peaks['hbm_gbs']}))
