"""ORACLE — TEST INFRASTRUCTURE ONLY.  Imports the *real* reference from /root/reference (build container only).

Used to pin the restatements in this directory and to mint tests/golden fixtures (tests/golden/make_golden.py).
Nothing that runs on the GPU box may call this: /root/reference does not exist there.

Bootstrap recipe = SURVEY.md Appendix A.5: alias numpy.float_ (removed in NumPy 2, used inside the njit bodies at
src/stereoimage_generation.py:138,177,197,223,229), stub absent third-party modules, put the reference on sys.path.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("DEPTHMAP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _stub(name):
    m = mock.MagicMock(name=name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__name__ = name
    m.__path__ = []
    sys.modules[name] = m
    return m


_done = False


def bootstrap():
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError("reference tree not present")
    import numpy
    if not hasattr(numpy, "float_"):
        numpy.float_ = numpy.float64
    for name in ["skimage", "skimage.measure", "timm", "timm.models", "timm.models.beit", "timm.models.layers",
                 "timm.models.registry", "diffusers", "diffusers.utils", "transformers", "matplotlib", "matplotlib.cm",
                 "gradio"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    d = sys.modules["diffusers"]
    if isinstance(d, mock.MagicMock):
        d.DiffusionPipeline = type("DiffusionPipeline", (), {})
        sys.modules["diffusers.utils"].BaseOutput = type("BaseOutput", (), {})
    t = sys.modules["timm.models.registry"]
    if isinstance(t, mock.MagicMock):
        t.register_model = lambda f: f
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _done = True


def stereo_module():
    bootstrap()
    import src.stereoimage_generation as m
    return m


def normalmap_module():
    bootstrap()
    import src.normalmap_generation as m
    return m


def dav2_class():
    bootstrap()
    from ddepth_anything_v2.depth_anything_v2.dpt import DepthAnythingV2
    return DepthAnythingV2
