"""ORACLE (test infrastructure only — the product never imports this): numpy / cv2 restatement of the BOOST driver, SURVEY.md
§8a row D9 ("Boosting Monocular Depth": resolution search, double estimation, patch selection, merge and blend).

Follows /root/reference/src/depthmap_generation.py:
  :774-941    estimateboost          whole-image double estimate, target resolution, patch loop (double estimate of the crop, merge
                                     with the base crop, degree-1 polyfit onto the base, Gaussian-mask blend), final cubic resize
  :944-953    generatemask           box of ones with a 15 % margin, Gaussian blur (sigma = size / 16), min-max normalised
  :956-966    rgb2gray, resizewithpool (skimage.measure.block_reduce(max): zero padded to a multiple of the block)
  :969-1024   calculateprocessingres the R_x search on a thresholded gradient map.  NOTE the reference passes the interpolation flag
                                     as cv2.resize's third POSITIONAL argument, which is `dst`: both resizes are bilinear
  :1028-1050  doubleestimate         low- and high-resolution estimate, both cubic-resized to 1024^2, merged, min-max normalised
  :1070-1177  generatepatchs, applyGridpatch, adaptiveselection, getGF_fromintegral
  :673-718    ImageandPatchs         (the live, second definition): rect and size scaled by `mergein_scale` and rounded
The two networks are parameters: `estimate(img, msize) -> float32 [h, w]` stands for singleestimate (:1053-1067, e.g.
oracle.leres.estimateleres) and `merge(outer, inner) -> float32 [1024, 1024] in (-1, 1)` for Pix2Pix4DepthModel.set_input + test +
fake_B (pix2pix/models/pix2pix4depth_model.py:96-116, oracle.pix2pix).  Pinned by tests/test_oracle_pin.py::test_boost_* against
the reference functions themselves with the same stand-in networks."""
from __future__ import annotations

import numpy as np

PIX2PIX_SIZE = 1024
R_THRESHOLD = 0.2
SCALE_THRESHOLD = 3


def receptive_field(model_type):
    """:777-786"""
    if model_type == 0:
        return 448
    if model_type == 1:
        return 512
    if model_type in (11, 12, 13, 14):
        return 518
    return 384


def generatemask(size):
    import cv2
    mask = np.zeros(size, dtype=np.float32)
    sigma = int(size[0] / 16)
    k = int(2 * np.ceil(2 * int(size[0] / 16)) + 1)
    my, mx = int(0.15 * size[0]), int(0.15 * size[1])
    mask[my:size[0] - my, mx:size[1] - mx] = 1
    mask = cv2.GaussianBlur(mask, (k, k), sigma)
    return ((mask - mask.min()) / (mask.max() - mask.min())).astype(np.float32)


def rgb2gray(rgb):
    return np.dot(rgb[..., :3], [0.2989, 0.5870, 0.1140])


def gradient_magnitude(gray):
    import cv2
    return np.abs(cv2.Sobel(gray, cv2.CV_64F, 0, 1, ksize=3)) + np.abs(cv2.Sobel(gray, cv2.CV_64F, 1, 0, ksize=3))


def block_reduce_max(img, n):
    """skimage.measure.block_reduce(img, (n, n), np.max): pad with zeros up to a multiple of n, max over n x n blocks."""
    h, w = img.shape
    ph, pw = (-h) % n, (-w) % n
    if ph or pw:
        img = np.pad(img, ((0, ph), (0, pw)), mode="constant", constant_values=0)
    H, W = img.shape
    return img.reshape(H // n, n, W // n, n).max(axis=(1, 3))


def calculateprocessingres(img, basesize, confidence=0.1, scale_threshold=3, whole_size_threshold=3000):
    import cv2
    speed_scale = 32
    image_dim = int(min(img.shape[0:2]))
    grad = gradient_magnitude(rgb2gray(img))
    grad = cv2.resize(grad, (image_dim, image_dim))                      # bilinear (see the module docstring)
    lo, hi = grad.min(), grad.max()
    middle = lo + 0.4 * (hi - lo)
    grad = (grad >= middle).astype(np.float64)
    k1 = np.ones((int(basesize / speed_scale),) * 2, float)
    k2 = np.ones((int(basesize / (4 * speed_scale)),) * 2, float)
    threshold = min(whole_size_threshold, scale_threshold * max(img.shape[:2]))
    outputsize_scale = basesize / speed_scale
    grad_resized = None
    for p_size in range(int(basesize / speed_scale), int(threshold / speed_scale), int(basesize / (2 * speed_scale))):
        n = int(np.floor(grad.shape[0] / p_size))
        grad_resized = cv2.resize(block_reduce_max(grad, n), (p_size, p_size))
        grad_resized = (grad_resized >= 0.5).astype(np.float64)
        dilated = cv2.dilate(grad_resized, k1, iterations=1)
        if (1 - dilated).mean() > confidence:
            break
        outputsize_scale = p_size
    patch_scale = cv2.dilate(grad_resized, k2, iterations=1).mean()
    return int(outputsize_scale * speed_scale), patch_scale


def _density(integral, rect):
    """getGF_fromintegral / area (:1167-1176); rect = [x, y, w, h]"""
    r1, r2, c1, c2 = rect[1], rect[1] + rect[3], rect[0], rect[0] + rect[2]
    return (integral[r2, c2] - integral[r1, c2] - integral[r2, c1] + integral[r1, c1]) / (rect[2] * rect[3])


def generatepatchs(img, base_size, factor):
    """-> [(key, {'rect': [x, y, w, h], 'size': w}), ...] sorted by size, largest first (stable, like the reference's sorted())."""
    import cv2
    grad = gradient_magnitude(rgb2gray(img))
    grad[grad < grad[grad > 0].mean()] = 0
    gf = grad.sum() / grad.size
    integral = cv2.integral(grad)
    blsize = int(round(base_size / 2))
    stride = int(round(blsize * 0.75))
    # applyGridpatch: square patches of side 2 * blsize on a stride grid, columns outermost
    grid = [[k - blsize, j - blsize, 2 * blsize, 2 * blsize]
            for k in range(blsize, img.shape[1] - blsize, stride) for j in range(blsize, img.shape[0] - blsize, stride)]
    # adaptiveselection: keep the patches at least as dense in gradients as the image, grow each while that stays true
    height, width = integral.shape
    step = int(32 / factor)
    chosen = []
    for bbox in grid:
        if _density(integral, bbox) < gf:
            continue
        trial = list(bbox)
        while True:
            trial = [trial[0] - int(step / 2), trial[1] - int(step / 2), trial[2] + step, trial[3] + step]
            if trial[0] < 0 or trial[1] < 0 or trial[1] + trial[3] >= height or trial[0] + trial[2] >= width:
                break
            if _density(integral, trial) < gf:
                break
            bbox = list(trial)
        chosen.append(bbox)
    items = [(str(i), {"rect": r, "size": r[2]}) for i, r in enumerate(chosen)]
    return sorted(items, key=lambda kv: kv[1]["size"], reverse=True)


def doubleestimate(img, size1, size2, estimate, merge):
    import cv2
    e1 = cv2.resize(estimate(img, size1), (PIX2PIX_SIZE, PIX2PIX_SIZE), interpolation=cv2.INTER_CUBIC)
    e2 = cv2.resize(estimate(img, size2), (PIX2PIX_SIZE, PIX2PIX_SIZE), interpolation=cv2.INTER_CUBIC)
    m = (np.asarray(merge(e1, e2), dtype=np.float32) + 1) / 2
    return (m - m.min()) / (m.max() - m.min())


def target_size(shape, whole_size, factor):
    """:822-830 -> (rows a, cols b) of the image the patches are cut from"""
    if shape[0] > shape[1]:
        a, b = 2 * whole_size, round(2 * whole_size * shape[1] / shape[0])
    else:
        a, b = round(2 * whole_size * shape[0] / shape[1]), 2 * whole_size
    return int(round(a / factor)), int(round(b / factor))


def scaled_rect(rect, size, scale):
    """ImageandPatchs.__getitem__ (:699-708)"""
    return np.round(np.array(rect) * scale).astype("int"), round(size * scale)


def estimateboost(img, model_type, estimate, merge, whole_size_threshold, info=None):
    """img: float [H, W, 3] in [0, 1] (what ModelHolder.get_raw_prediction hands over, :381) -> float [H, W]"""
    import cv2
    rf = receptive_field(model_type)
    patch_netsize = 2 * rf
    mask_org = generatemask((3000, 3000))
    mask = mask_org.copy()
    input_resolution = img.shape
    whole_size, patch_scale = calculateprocessingres(img, rf, R_THRESHOLD, SCALE_THRESHOLD, whole_size_threshold)
    whole_estimate = doubleestimate(img, rf, whole_size, estimate, merge)
    factor = max(min(1, 4 * patch_scale * whole_size / whole_size_threshold), 0.2)
    a, b = target_size(img.shape, whole_size, factor)
    img = cv2.resize(img, (b, a), interpolation=cv2.INTER_CUBIC)
    patchset = generatepatchs(img, rf * 2, factor)
    mergein_scale = input_resolution[0] / img.shape[0]
    rgb = cv2.resize(img, (round(img.shape[1] * mergein_scale), round(img.shape[0] * mergein_scale)), interpolation=cv2.INTER_CUBIC)
    base = cv2.resize(whole_estimate, (round(img.shape[1] * mergein_scale), round(img.shape[0] * mergein_scale)),
                      interpolation=cv2.INTER_CUBIC)
    updated = base.copy()
    if info is not None:
        info.update(whole_size=whole_size, patch_scale=patch_scale, factor=factor, target=(a, b), patches=[kv[1]["rect"] for kv in patchset])
    for _, entry in patchset:
        rect, _ = scaled_rect(entry["rect"], entry["size"], mergein_scale)
        x1, y1, x2, y2 = rect[0], rect[1], rect[0] + rect[2], rect[1] + rect[3]
        patch_rgb = rgb[y1:y2, x1:x2]
        patch_base = base[y1:y2, x1:x2]
        org_size = patch_base.shape
        est = doubleestimate(patch_rgb, rf, patch_netsize, estimate, merge)
        est = cv2.resize(est, (PIX2PIX_SIZE, PIX2PIX_SIZE), interpolation=cv2.INTER_CUBIC)
        patch_base = cv2.resize(patch_base, (PIX2PIX_SIZE, PIX2PIX_SIZE), interpolation=cv2.INTER_CUBIC)
        mapped = (np.asarray(merge(patch_base, est), dtype=np.float32) + 1) / 2
        coef = np.polyfit(mapped.reshape(-1), patch_base.reshape(-1), deg=1)
        merged = np.polyval(coef, mapped.reshape(-1)).reshape(mapped.shape)
        merged = cv2.resize(merged, (org_size[1], org_size[0]), interpolation=cv2.INTER_CUBIC)
        if mask.shape != org_size:
            mask = cv2.resize(mask_org, (org_size[1], org_size[0]), interpolation=cv2.INTER_LINEAR)
        updated[y1:y2, x1:x2] = updated[y1:y2, x1:x2] * (1 - mask) + merged * mask
    return cv2.resize(updated, (input_resolution[1], input_resolution[0]), interpolation=cv2.INTER_CUBIC)
