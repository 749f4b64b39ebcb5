"""Importable alias of the product package.

The product lives in ``stable-diffusion-webui-depthmap-script_b200/`` (a directory name Python cannot import because of
the hyphens); this shim makes it importable as ``depthmap_b200`` by pointing ``__path__`` at that directory and
executing its ``__init__.py`` in this module's namespace.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "stable-diffusion-webui-depthmap-script_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
