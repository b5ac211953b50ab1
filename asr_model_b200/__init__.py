"""Import shim: the package directory is named ``asr-model_b200`` (not a Python
identifier), so ``import asr_model_b200`` resolves here and this module re-points
its search path at the real directory and runs its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "asr-model_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
