"""Import shim: the product package lives in ``triton-racer-sim_b200/`` (the directory name the project
layout prescribes), which is not a valid Python identifier.  ``import triton_racer_sim_b200`` resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "triton-racer-sim_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
