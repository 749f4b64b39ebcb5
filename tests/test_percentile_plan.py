"""Host logic of the "Outliers" clip: core.percentile_plan must reproduce where np.percentile (method "linear") reads and the
float64 weight it interpolates with — checked by rebuilding np.percentile's result from the plan with numpy's _lerp formula."""
import numpy as np
import pytest

from depthmap_b200.core import percentile_plan


def _lerp(a, b, t):      # numpy/lib/_function_base_impl.py:_lerp for float32 a, b and float64 t
    diff = np.float32(b) - np.float32(a)
    return np.float64(b) - np.float64(diff) * (1 - t) if t >= 0.5 else np.float64(a) + np.float64(diff) * t


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000, 518 * 518, 512 * 512 + 1])
def test_plan_rebuilds_np_percentile(n):
    rng = np.random.default_rng(n)
    a = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 3)).astype(np.float32)
    srt = np.sort(a)
    fracs = [0.0, 1.0, 0.05, 0.95, 0.5, 1 / 3, 0.123456789, 0.999999, 1e-9]
    want = np.percentile(a, [f * 100.0 for f in fracs])
    assert want.dtype == np.float64
    for f, w, (p, x, g) in zip(fracs, want, percentile_plan(n, fracs)):
        assert 0 <= p < n and 0 <= x < n
        got = _lerp(srt[p], srt[x], g)
        assert got == w or (np.isnan(got) and np.isnan(w)), (n, f, p, x, g, got, w)


def test_plan_rejects_out_of_range():
    with pytest.raises(ValueError):
        percentile_plan(10, [1.5])
