"""Seeded synthetic inputs: blurred-noise RGB and blurred-noise u16 depth."""
from __future__ import annotations

import numpy as np


def _blur_noise(rng, h, w, sigma):
    import cv2
    x = rng.random((h, w)).astype(np.float64)
    k = int(max(3, (int(sigma * 3) | 1)))
    x = cv2.GaussianBlur(x, (k, k), sigma, borderType=cv2.BORDER_REFLECT_101)
    x -= x.min()
    m = x.max()
    return x / m if m > 0 else x


def synth_rgb(h, w, seed=0):
    rng = np.random.default_rng(1234 + seed)
    chans = [_blur_noise(rng, h, w, max(1.0, h / 64.0)) for _ in range(3)]
    img = np.stack(chans, axis=2) * 255.0
    img = 0.95 * img + 0.05 * rng.random((h, w, 3)) * 255.0
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_depth_u16(h, w, seed=0, edges=True):
    rng = np.random.default_rng(4321 + seed)
    d = _blur_noise(rng, h, w, max(1.0, h / 32.0))
    if edges:  # a few hard depth discontinuities so gap-fill paths are exercised
        for _ in range(4):
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
            y1, x1 = min(h, y0 + int(rng.integers(2, max(3, h // 3)))), min(w, x0 + int(rng.integers(2, max(3, w // 3))))
            d[y0:y1, x0:x1] = np.clip(d[y0:y1, x0:x1] + rng.uniform(-0.5, 0.5), 0, 1)
        d -= d.min()
        d /= max(d.max(), 1e-12)
    return np.clip(d * 65535.0 + 0.5, 0, 65535).astype(np.uint16)


def noise_rgb(h, w, seed=0):
    return np.random.default_rng(99 + seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def noise_depth_u16(h, w, seed=0):
    return np.random.default_rng(77 + seed).integers(0, 65536, (h, w), dtype=np.uint16)
