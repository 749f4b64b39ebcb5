"""ORACLE — TEST INFRASTRUCTURE ONLY.  Python face of ``stereo_oracle.c`` with the reference's signatures.

Restates /root/reference/src/stereoimage_generation.py:13-92 (mode handling, balance split, px conversion); the
per-row algorithms live in C (``oracle_stereo_naive`` = :95-159, ``oracle_stereo_polylines`` = :162-283).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
from PIL import Image

from . import lib

_FILL_NAIVE = {"none": 0, "naive": 1, "naive_interpolating": 2}
_FILL_POLY = {"polylines_soft": 0, "polylines_sharp": 1}


def apply_stereo_divergence(original_image, depth, divergence, separation, stereo_offset_exponent, fill_technique,
                            nthreads=None):
    """stereoimage_generation.py:77-92"""
    original_image = np.ascontiguousarray(np.asarray(original_image), dtype=np.uint8)
    depth = np.asarray(depth)
    assert original_image.shape[:2] == depth.shape, 'Depthmap and the image must have the same size'
    depth_min = depth.min()
    depth_max = depth.max()
    with np.errstate(all="ignore"):
        normalized_depth = (depth - depth_min) / (depth_max - depth_min)
    normalized_depth = np.ascontiguousarray(normalized_depth, dtype=np.float64)
    divergence_px = (divergence / 100.0) * original_image.shape[1]
    separation_px = (separation / 100.0) * original_image.shape[1]
    h, w, c = original_image.shape
    assert c == 3
    out = np.empty_like(original_image)
    L = lib()
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    if fill_technique in _FILL_NAIVE:
        rc = L.oracle_stereo_naive(original_image.ctypes.data, normalized_depth.ctypes.data, h, w,
                                   float(divergence_px), float(separation_px), float(stereo_offset_exponent),
                                   _FILL_NAIVE[fill_technique], out.ctypes.data)
    elif fill_technique in _FILL_POLY:
        rc = L.oracle_stereo_polylines(original_image.ctypes.data, normalized_depth.ctypes.data, h, w,
                                       float(divergence_px), float(separation_px), float(stereo_offset_exponent),
                                       _FILL_POLY[fill_technique], int(nthreads), out.ctypes.data)
    else:
        return None  # the reference falls through (:85-92)
    if rc != 0:
        raise RuntimeError(f"oracle stereo failed rc={rc}")
    return out


def overlap_red_cyan(im1, im2):
    """stereoimage_generation.py:286-307"""
    im1 = np.ascontiguousarray(im1, dtype=np.uint8)
    im2 = np.ascontiguousarray(im2, dtype=np.uint8)
    out = np.empty_like(im2)
    lib().oracle_overlap_red_cyan(im1.ctypes.data, im2.ctypes.data, im2.shape[0], im2.shape[1], out.ctypes.data)
    return out


def create_stereoimages(original_image, depthmap, divergence, separation=0.0, modes=None,
                        stereo_balance=0.0, stereo_offset_exponent=1.0, fill_technique='polylines_sharp',
                        return_arrays=False, nthreads=None):
    """stereoimage_generation.py:13-74"""
    if modes is None:
        modes = ['left-right']
    if not isinstance(modes, list):
        modes = [modes]
    if len(modes) == 0:
        return []
    original_image = np.asarray(original_image)
    balance = (stereo_balance + 1) / 2
    left_eye = original_image if balance < 0.001 else \
        apply_stereo_divergence(original_image, depthmap, +1 * divergence * balance, -1 * separation,
                                stereo_offset_exponent, fill_technique, nthreads)
    right_eye = original_image if balance > 0.999 else \
        apply_stereo_divergence(original_image, depthmap, -1 * divergence * (1 - balance), separation,
                                stereo_offset_exponent, fill_technique, nthreads)
    results = []
    for mode in modes:
        if mode == 'left-right':
            results.append(np.hstack([left_eye, right_eye]))
        elif mode == 'right-left':
            results.append(np.hstack([right_eye, left_eye]))
        elif mode == 'top-bottom':
            results.append(np.vstack([left_eye, right_eye]))
        elif mode == 'bottom-top':
            results.append(np.vstack([right_eye, left_eye]))
        elif mode == 'red-cyan-anaglyph':
            results.append(overlap_red_cyan(left_eye, right_eye))
        elif mode == 'left-only':
            results.append(left_eye)
        elif mode == 'only-right':
            results.append(right_eye)
        elif mode == 'cyan-red-reverseanaglyph':
            results.append(overlap_red_cyan(right_eye, left_eye))
        else:
            raise Exception('Unknown mode')
    if return_arrays:
        return results
    return [Image.fromarray(r) for r in results]
