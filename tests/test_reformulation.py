"""CPU: the data-parallel reformulation that csrc/stereo.cu implements (numpy model) equals the sequential oracle."""
import numpy as np
import pytest

from oracle import stereo as ost
from reformulation_model import stereo_eye
from synth import noise_depth_u16, noise_rgb, synth_depth_u16, synth_rgb

CASES = [(1.25, 0.0, 1.0), (-1.25, 0.0, 1.0), (5.0, -1.0, 1.0), (-7.5, 2.0, 2.0), (0.025, 0.0, 1.0), (-1.0, 0.0, 1.0)]


@pytest.mark.parametrize("fill", ['none', 'naive', 'naive_interpolating', 'polylines_soft', 'polylines_sharp'])
def test_model_equals_oracle(fill):
    for (h, w, seed, kind) in [(6, 53, 0, 'smooth'), (6, 64, 1, 'noise'), (4, 200, 4, 'noise')]:
        img = synth_rgb(h, w, seed) if kind == 'smooth' else noise_rgb(h, w, seed)
        dep = synth_depth_u16(h, w, seed) if kind == 'smooth' else noise_depth_u16(h, w, seed)
        if seed == 1:
            img[1:4, 10:30] = 0
        for divp, sepp, ex in CASES:
            a = ost.apply_stereo_divergence(img, dep, divp, sepp, ex, fill)
            b = stereo_eye(img, dep, divp / 100.0 * w, sepp / 100.0 * w, ex, fill)
            assert np.array_equal(a, b), (fill, h, w, divp, sepp, ex)
