"""Import alias for the package directory `kitamr.jl_b200/` (whose name, fixed by the project
layout, is not a valid Python identifier).  `import kitamr_jl_b200` resolves every submodule
from that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kitamr.jl_b200"))

from . import abi  # noqa: E402,F401
