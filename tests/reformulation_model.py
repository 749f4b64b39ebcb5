"""numpy model of the data-parallel reformulation used by csrc/stereo.cu (same steps, same arithmetic order).

It exists so the *algorithmic* claims behind the CUDA kernel (prefix-max winner search, bucket sort, closed-form
naive_interpolating fill) are checked against the oracle on CPU, independent of any GPU.  Not part of the product.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-7


def _vertex_x(ndp_row, w, div_px, sep_px, sharp):
    col = np.arange(w, dtype=np.float64)
    coord_d = ndp_row * div_px
    coord_x = col + 0.5 + coord_d + sep_px
    if sharp:
        x = np.empty(2 * w + 2)
        x[1:-1:2] = coord_x - 0.45
        x[2:-1:2] = coord_x + 0.45
        src = np.empty(2 * w + 2, dtype=np.int64)
        src[1:-1:2] = np.arange(w)
        src[2:-1:2] = np.arange(w)
    else:
        x = np.empty(w + 2)
        x[1:-1] = coord_x
        src = np.empty(w + 2, dtype=np.int64)
        src[1:-1] = np.arange(w)
    x[0] = -1.0 * w
    x[-1] = 2.0 * w
    src[0] = 0
    src[-1] = w - 1
    return x, src


def polylines_row(img_row, ndp_row, div_px, sep_px, sharp):
    w = img_row.shape[0]
    X, src = _vertex_x(ndp_row, w, div_px, sep_px, sharp)
    n = X.shape[0]
    fwd = not (div_px < 0)
    pm = np.maximum.accumulate(X) if fwd else np.minimum.accumulate(X[::-1])[::-1]
    # stable order of vertices 0..n-2 by (x, index); sentinel n-1 stays last
    order = np.concatenate([np.lexsort((np.arange(n - 1), X[:n - 1])), [n - 1]])
    xs = X[order]
    out = np.zeros((w, 3), dtype=np.uint8)
    imgf = img_row.astype(np.float64)
    for col in range(w):
        color = np.full(3, 0.5)
        i = int(np.searchsorted(xs[:n - 1], col, side='left')) - 1  # last vertex with x < col
        xi = xs[i]
        while xi < col + 1:
            xn = xs[i + 1]
            cf = (xi if xi > col else float(col)) + EPS
            ct = (xn if xn < col + 1 else float(col + 1)) - EPS
            sig = ct - cf
            ctr = cf + 0.5 * sig
            # fwd: (first t with pm[t] >= ctr) - 1 ; bwd: last t with pm[t] < ctr -- both are searchsorted_left - 1
            s = int(np.searchsorted(pm, ctr, side='left')) - 1
            s = max(0, min(s, n - 2))
            cl, cr = src[s], src[s + 1]
            if cl == cr:
                color = color + imgf[cl] * sig
            else:
                k = (ctr - X[s]) / (X[s + 1] - X[s])
                color = color + (imgf[cl] * (1.0 - k) + imgf[cr] * k) * sig
            i += 1
            xi = xn
        out[col] = color.astype(np.int64).astype(np.uint8)
    return out


def naive_row(img_row, ndp_row, div_px, sep_px, fill):
    w = img_row.shape[0]
    v = ndp_row * div_px + sep_px
    cd = np.arange(w) + np.trunc(v).astype(np.int64)
    ok = (cd >= 0) & (cd < w)
    take_max = div_px < 0
    winner = np.full(w, -1 if take_max else 2 ** 31 - 1, dtype=np.int64)
    if take_max:
        np.maximum.at(winner, cd[ok], np.arange(w)[ok])
        filled = winner >= 0
    else:
        np.minimum.at(winner, cd[ok], np.arange(w)[ok])
        filled = winner != 2 ** 31 - 1
    scat = np.zeros((w, 3), dtype=np.uint8)
    scat[filled] = img_row[winner[filled]]
    if fill == 'none':
        return scat
    if fill == 'naive':
        out = scat.copy()
        lim = abs(int(div_px))
        for c in np.where(~filled)[0]:
            for o in range(1, lim + 2):
                if c + o < w and filled[c + o]:
                    out[c] = scat[c + o]
                    break
                if c - o >= 0 and filled[c - o]:
                    out[c] = scat[c - o]
                    break
        return out
    # naive_interpolating, closed form
    nonblack = scat.astype(np.int64).sum(axis=1) != 0
    V = filled & nonblack
    idx = np.arange(w)
    lastV = np.maximum.accumulate(np.where(V, idx, -1))
    nextV = np.minimum.accumulate(np.where(V, idx, w)[::-1])[::-1]
    nextU = np.minimum.accumulate(np.where(~filled, idx, w)[::-1])[::-1]
    out = scat.copy()
    for p in range(w):
        v_ = lastV[p]
        l = nextU[v_ + 1] if v_ + 1 < w else w
        if l <= p:
            r = nextV[p]
            lb = scat[l - 1].astype(np.int64) if l > 0 else np.zeros(3, np.int64)
            rb = scat[r].astype(np.int64) if r < w else np.zeros(3, np.int64)
            if lb.sum() == 0:
                lb = rb
            elif rb.sum() == 0:
                rb = lb
            step = (rb.astype(np.float64) - lb) / (1 + r - l)
            inc = np.trunc(step * (p - l + 1)).astype(np.int64)
            out[p] = ((lb + inc) % 256).astype(np.uint8)
    return out


def stereo_eye(img, depth_u16, div_px, sep_px, exponent, fill):
    d = depth_u16.astype(np.uint16)
    mn, mx = int(d.min()), int(d.max())
    nd = (d.astype(np.float64) - mn) / float(mx - mn)
    ndp = nd if exponent == 1.0 else (nd * nd if exponent == 2.0 else nd ** exponent)
    h, w = d.shape
    out = np.zeros_like(img)
    for y in range(h):
        if fill in ('polylines_sharp', 'polylines_soft'):
            out[y] = polylines_row(img[y], ndp[y], div_px, sep_px, fill == 'polylines_sharp')
        else:
            out[y] = naive_row(img[y], ndp[y], div_px, sep_px, fill)
    return out
