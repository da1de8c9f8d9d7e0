"""Import alias for the package directory `gbx-lm_b200/`.

The project layout fixes the package directory name as `gbx-lm_b200/`; a hyphen cannot appear in a
Python import statement, so this stub makes `import gbx_lm_b200` resolve to that directory: its
`__path__` points there and the real `__init__.py` is executed in this module's namespace."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gbx-lm_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _os, _real
