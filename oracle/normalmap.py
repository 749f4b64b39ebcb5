"""ORACLE — TEST INFRASTRUCTURE ONLY.  Python faces of ``normalmap_oracle.c``.

* :func:`create_normalmap`  — /root/reference/src/normalmap_generation.py:5-56
* :func:`normalize_to_u16`  — /root/reference/src/core.py:189-211 (+ convert_to_i16 :44-50), model-prediction branch
"""
from __future__ import annotations

import numpy as np
from PIL import Image

from . import lib


def create_normalmap(depthmap, pre_blur=None, sobel_gradient=3, post_blur=None, invert=False, return_array=False):
    d = np.ascontiguousarray(np.asarray(depthmap), dtype=np.uint16)
    h, w = d.shape
    out = np.empty((h, w, 3), dtype=np.uint8)
    rc = lib().oracle_normalmap(d.ctypes.data, h, w,
                                int(pre_blur) if pre_blur is not None and pre_blur > 0 else 0,
                                int(sobel_gradient) if sobel_gradient is not None and sobel_gradient > 0 else 0,
                                int(post_blur) if post_blur is not None and post_blur > 0 else 0,
                                1 if invert else 0, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle normalmap failed rc={rc}")
    return out if return_array else Image.fromarray(out)


def normalize_to_u16(raw_prediction, invert=False, clipdepth=False, clipdepth_mode="Range", far=0.0, near=1.0):
    """Funnel post-processing of one model prediction -> uint16 depth (near = bright)."""
    p = np.ascontiguousarray(raw_prediction, dtype=np.float32)
    out = np.empty(p.shape, dtype=np.uint16)
    if clipdepth and clipdepth_mode == "Outliers":
        # core.py:200-202 — numpy percentile path (kept in numpy: np.percentile's float64 linear interpolation)
        if not abs(p.max() - p.min()) > np.finfo("float").eps:
            return np.zeros(p.shape, dtype=np.uint16)
        o = np.copy(p)
        if invert:
            o *= -1
        fb, nb = np.percentile(o, [far * 100.0, near * 100.0])
        o = np.clip(o, fb, nb)
        o = (o - o.min()) / (o.max() - o.min())
        return np.clip(o * 65536 + 0.0001, 0, 65536 - 0.1).astype("uint16")
    mode = 1 if (clipdepth and clipdepth_mode == "Range") else 0
    lib().oracle_normalize_u16(p.ctypes.data, p.size, 1 if invert else 0, mode, float(far), float(near), out.ctypes.data)
    return out
