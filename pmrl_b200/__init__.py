"""Import shim: the package lives in `pm-rl_b200/` (not an importable name); this maps `pmrl_b200` onto it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pm-rl_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
