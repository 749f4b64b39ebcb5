"""CPU: the host side of BOOST — the control plane of depthmap_b200.boost against the oracle restatement
(which tests/test_oracle_pin.py pins to the reference functions): resolution search, patch selection, scaled crops, the separable
Gaussian mask profile, and the round-robin patch schedule of the patch-parallel mode (§8e)."""
import numpy as np
import pytest

from synth import synth_rgb


def _img(h, w, seed):
    import cv2
    return cv2.cvtColor(synth_rgb(h, w, seed), cv2.COLOR_BGR2RGB) / 255.0


@pytest.mark.parametrize("hw,rmax", [((300, 420), 1600), ((520, 360), 1200), ((256, 256), 900)])
def test_plan_equals_oracle(hw, rmax):
    from depthmap_b200 import boost
    from oracle import boost as ob
    img = _img(hw[0], hw[1], 11)
    p = boost.plan(img, 0, rmax)
    whole, patch_scale = ob.calculateprocessingres(img, 448, 0.2, 3, rmax)
    assert (p["whole"], p["patch_scale"]) == (whole, patch_scale)
    factor = max(min(1, 4 * patch_scale * whole / rmax), 0.2)
    a, b = ob.target_size(img.shape, whole, factor)
    assert p["target"] == (a, b)
    import cv2
    big = cv2.resize(img, (b, a), interpolation=cv2.INTER_CUBIC)
    want = ob.generatepatchs(big, 896, factor)
    assert [kv[1]["rect"] for kv in want] == p["rects"]
    scale = hw[0] / a
    assert p["work"] == (round(a * scale), round(b * scale))
    for kv, got in zip(want, p["scaled_rects"]):
        rect, _ = ob.scaled_rect(kv[1]["rect"], kv[1]["size"], scale)
        crop = np.zeros(p["work"])[rect[1]:rect[1] + rect[3], rect[0]:rect[0] + rect[2]]       # numpy clips like the reference's impatch
        assert got[:2] == [rect[0], rect[1]] and (got[3], got[2]) == crop.shape


def test_mask_profile_is_the_separable_form_of_generatemask():
    from depthmap_b200 import boost
    from oracle import boost as ob
    prof = boost.mask_profile(3000)
    mask = ob.generatemask((3000, 3000))
    assert prof.shape == (3000,) and prof.max() == 1.0 and prof.min() == 0.0
    for r in (0, 449, 600, 1500, 2399, 2999):
        assert np.abs(mask[r] - prof[r] * prof).max() < 2e-6
    assert np.abs(mask[:, 777] - prof * prof[777]).max() < 2e-6


def test_receptive_fields():
    from depthmap_b200 import boost
    from oracle import boost as ob
    for t in range(15):
        assert boost.receptive_field(t) == ob.receptive_field(t)


def test_plan_without_patches():
    """an image smaller than the patch grid selects nothing: estimateboost then returns the resized whole-image double estimate"""
    from depthmap_b200 import boost
    from oracle import boost as ob
    img = _img(200, 200, 3)
    p = boost.plan(img, 0, 1600)
    assert p["rects"] == [] and p["scaled_rects"] == []
    whole, patch_scale = ob.calculateprocessingres(img, 448, 0.2, 3, 1600)
    assert p["whole"] == whole
