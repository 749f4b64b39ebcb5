"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path, used solely as the parity checker by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.
The product package (``stable-diffusion-webui-depthmap-script_b200`` a.k.a. ``depthmap_b200``) never imports this.

Pieces
------
* ``stereo_oracle.c`` / ``normalmap_oracle.c`` — plain C, compiled by :func:`build` with ``gcc -O2 -ffp-contract=off``
  into ``oracle/_build/liboracle.so`` (git-ignored, travels to the GPU box).
* :mod:`oracle.stereo`, :mod:`oracle.normalmap`, :mod:`oracle.postprocess` — Python faces with the reference's
  signatures (``create_stereoimages``, ``create_normalmap``, funnel normalise + ``convert_to_i16``).
* :mod:`oracle.dav2` / :mod:`oracle.beit_dpt` — plain-PyTorch fp32 functional restatements of the depth networks.
* :mod:`oracle.ref_loader` — imports the *real* reference from ``/root/reference`` (build container only) to pin
  the restatements and to mint ``tests/golden`` fixtures.

Pinning status: stereo / normal map / normalise / Depth-Anything-V2 are pinned against the reference's own code run in
the build container (``tests/test_oracle_pin.py`` + committed goldens).  DPT-BEiT (timm absent) is "parity unpinned".
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "liboracle.so")
_SRCS = ["stereo_oracle.c", "normalmap_oracle.c"]
_lock = threading.Lock()
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle (idempotent). Returns the path of the shared object."""
    os.makedirs(_BUILD, exist_ok=True)
    srcs = [os.path.join(_HERE, s) for s in _SRCS]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", _SO] + srcs + ["-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            _lib = ctypes.CDLL(build())
            c = ctypes
            _lib.oracle_normalize_depth_u16.argtypes = [c.c_void_p, c.c_int64, c.c_void_p]
            _lib.oracle_normalize_depth_u16.restype = None
            _lib.oracle_stereo_naive.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_double, c.c_double,
                                                 c.c_double, c.c_int, c.c_void_p]
            _lib.oracle_stereo_naive.restype = c.c_int
            _lib.oracle_stereo_polylines.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_double, c.c_double,
                                                     c.c_double, c.c_int, c.c_int, c.c_void_p]
            _lib.oracle_stereo_polylines.restype = c.c_int
            _lib.oracle_overlap_red_cyan.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_void_p]
            _lib.oracle_overlap_red_cyan.restype = None
            _lib.oracle_num_threads.restype = c.c_int
            _lib.oracle_normalmap.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
            _lib.oracle_normalmap.restype = c.c_int
            _lib.oracle_normalize_u16.argtypes = [c.c_void_p, c.c_int64, c.c_int, c.c_int, c.c_float, c.c_float, c.c_void_p]
            _lib.oracle_normalize_u16.restype = c.c_int
        return _lib
