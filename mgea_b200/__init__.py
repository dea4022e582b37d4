"""Import shim: ``import mgea_b200`` loads the package in ``music-generation-emotion-adaptive_b200/``.

The product directory carries the reference repository's name, which is not a valid Python
identifier; this shim points the package search path at it and runs its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "music-generation-emotion-adaptive_b200")
if not _os.path.isdir(_real):
    raise ImportError(f"product package directory missing: {_real}")
__path__[:] = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
