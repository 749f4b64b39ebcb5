"""CPU: the C oracle reproduces the committed golden vectors (minted from the real reference by golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import normalmap as onm
from oracle import stereo as ost

G = os.path.join(os.path.dirname(__file__), "golden")
sys_path_ok = True
from golden.make_golden import FILLS, NORMAL_CASES, STEREO_CASES  # noqa: E402

MODES = ['left-right', 'red-cyan-anaglyph', 'top-bottom', 'cyan-red-reverseanaglyph']


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "stereo_normal_golden.npz"))


@pytest.mark.parametrize("name", ["smooth", "noise", "black", "flat"])
@pytest.mark.parametrize("fill", FILLS)
def test_oracle_stereo_matches_golden(gold, name, fill):
    img, dep = gold[f"in_{name}_rgb"], gold[f"in_{name}_depth"]
    for ci, (div, sep, bal, ex) in enumerate(STEREO_CASES):
        with np.errstate(all="ignore"):
            res = ost.create_stereoimages(img, dep, div, sep, MODES, bal, ex, fill, return_arrays=True)
        for m, r in zip(MODES, res):
            assert np.array_equal(r, gold[f"stereo_{name}_{fill}_{ci}_{m}"]), (name, fill, ci, m)


@pytest.mark.parametrize("name", ["smooth", "noise", "black"])
def test_oracle_normalmap_matches_golden(gold, name):
    dep = gold[f"in_{name}_depth"]
    for ci, (pb, sb, qb, inv) in enumerate(NORMAL_CASES):
        r = onm.create_normalmap(dep, pb, sb, qb, inv, return_array=True)
        assert np.array_equal(r, gold[f"normal_{name}_{ci}"]), (name, ci)


def test_oracle_normalize_matches_golden():
    g = np.load(os.path.join(G, "normalize_golden.npz"))
    for i in range(6):
        raw = g[f"pred_{i}"]
        for inv in (False, True):
            for ci, (clip, far, near) in enumerate([(False, 0.0, 1.0), (True, 0.1, 0.8), (True, 0.0, 0.5)]):
                r = onm.normalize_to_u16(raw, inv, clip, "Range", far, near)
                assert np.array_equal(r, g[f"u16_{i}_{int(inv)}_{ci}"]), (i, inv, ci)


def test_oracle_degenerate_prediction_is_black():
    assert onm.normalize_to_u16(np.full((5, 7), 3.25, np.float32)).sum() == 0


def test_oracle_stereo_api_edges():
    img = np.zeros((4, 6, 3), np.uint8)
    dep = np.arange(24, dtype=np.uint16).reshape(4, 6)
    assert ost.create_stereoimages(img, dep, 2.5, modes=[]) == []
    with pytest.raises(Exception, match="Unknown mode"):
        ost.create_stereoimages(img, dep, 2.5, modes=["sideways"])
    with pytest.raises(AssertionError):
        ost.create_stereoimages(img, dep[:3], 2.5)
    out = ost.create_stereoimages(img, dep, 2.5, modes="left-right")
    assert out[0].size == (12, 4)
