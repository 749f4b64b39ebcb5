"""ORACLE (test infrastructure only — the product never imports this): numpy restatement of the cross-frame normalisation
of video mode, SURVEY.md §8(f) rank 1.  Follows /root/reference/src/video_mode.py:103-128 (`process_predicitons`).

  smoothening 'none'          every frame is scaled with the GLOBAL minimum / maximum over all frames
  smoothening 'experimental'  a 5-tap temporal blend (0.10, 0.20, 0.40, 0.20, 0.10, indices clamped at the ends) is used
                              only to pick the scaling bounds: the 0.5 / 99.5 percentiles of the blended stack; the frames that
                              are scaled are the ORIGINAL ones, and the result is not clipped to [0, 1]
  anything else               predictions returned untouched
Dtype flow is numpy's: float32 frames stay float32 under 'none' (min / max are float32 scalars of the same dtype); under
'experimental' np.percentile returns float64 scalars, which promote every scaled frame to float64.
Pinned to the reference function by tests/test_oracle_pin.py::test_video_normalisation_equals_reference.  Round 1 ships only
this oracle; the multi-GPU form (one all-reduce of min / max, a distributed percentile) is round-2 work."""
from __future__ import annotations

import numpy as np

TAPS = (0.10, 0.20, 0.40, 0.20, 0.10)


def _global_scaling(frames, lo=None, hi=None):
    lo = lo if lo is not None else min(f.min() for f in frames)
    hi = hi if hi is not None else max(f.max() for f in frames)
    return [(f - lo) / (hi - lo) for f in frames]


def process_predictions(predictions, smoothening="none"):
    if smoothening == "none":
        return _global_scaling(predictions)
    if smoothening == "experimental":
        n = len(predictions)
        blended = []
        for i in range(n):
            acc = np.zeros_like(predictions[i])
            for u, mul in enumerate(TAPS):
                acc += mul * predictions[min(max(0, i + u - 2), n - 1)]
            blended.append(acc)
        lo, hi = np.percentile(np.stack(blended), [0.5, 99.5])
        return _global_scaling(predictions, lo, hi)
    return predictions
