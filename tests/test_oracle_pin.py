"""CPU, build container only: pins the oracle to the REAL reference imported from /root/reference.

Skipped where the reference tree is absent (the GPU box); there the committed goldens carry the pin."""
import warnings

import numpy as np
import pytest

from oracle import normalmap as onm
from oracle import ref_loader
from oracle import stereo as ost
from synth import noise_depth_u16, noise_rgb, synth_depth_u16, synth_rgb

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    warnings.filterwarnings("ignore")
    return ref_loader.stereo_module(), ref_loader.normalmap_module()


@pytest.mark.parametrize("fill", ['none', 'naive', 'naive_interpolating', 'polylines_soft', 'polylines_sharp'])
def test_stereo_oracle_equals_reference(ref, fill):
    rs, _ = ref
    rng = np.random.default_rng(11)
    for (h, w, seed, kind) in [(21, 67, 10, 'smooth'), (16, 50, 11, 'noise')]:
        img = synth_rgb(h, w, seed) if kind == 'smooth' else noise_rgb(h, w, seed)
        dep = synth_depth_u16(h, w, seed) if kind == 'smooth' else noise_depth_u16(h, w, seed)
        for _ in range(3):
            div, sep = float(rng.uniform(0.05, 15)), float(rng.uniform(-5, 5))
            bal, ex = float(rng.uniform(-1, 1)), float(rng.choice([1.0, 2.0]))
            a = rs.create_stereoimages(img, dep, div, sep, ['left-right', 'red-cyan-anaglyph'], bal, ex, fill)
            b = ost.create_stereoimages(img, dep, div, sep, ['left-right', 'red-cyan-anaglyph'], bal, ex, fill)
            for x, y in zip(a, b):
                assert np.array_equal(np.asarray(x), np.asarray(y)), (fill, div, sep, bal, ex)


def test_stereo_oracle_float_depth_equals_reference(ref):
    rs, _ = ref
    img = noise_rgb(12, 40, 5)
    dep = np.random.default_rng(5).random((12, 40)).astype(np.float32)
    for fill in ['naive', 'polylines_sharp']:
        a = rs.create_stereoimages(img, dep, 3.0, fill_technique=fill)[0]
        b = ost.create_stereoimages(img, dep, 3.0, fill_technique=fill)[0]
        assert np.array_equal(np.asarray(a), np.asarray(b))


def test_normalmap_oracle_equals_reference(ref):
    _, rn = ref
    for seed in range(3):
        dep = synth_depth_u16(30, 41, seed) if seed else noise_depth_u16(30, 41, seed)
        for (pb, sb, qb, inv) in [(None, 3, None, False), (None, 5, None, True), (None, None, None, False),
                                  (3, 3, 3, False), (None, 11, None, False), (9, 3, None, True), (None, 3, 11, False),
                                  (None, 31, None, False), (31, 3, 31, False)]:
            a = np.asarray(rn.create_normalmap(dep, pb, sb, qb, inv))
            b = onm.create_normalmap(dep, pb, sb, qb, inv, return_array=True)
            assert np.array_equal(a, b), (seed, pb, sb, qb, inv)
