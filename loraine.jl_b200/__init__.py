"""loraine.jl_b200 -- Python host side of the B200-native replacement of Loraine.jl's interior-point hot path.

The directory name is not an importable identifier; load it with `__graft_entry__.load_package()` (or
importlib, see that function).  The host mirrors the reference's `Solvers` module interface (same function names,
argument meaning and error behaviour) and drives the C-ABI library `libloraine_b200.so`; there is no CPU fallback:
every numerical step of the hot path is a call into the CUDA library, and importing `.solver` raises when the
library cannot be loaded.
"""
from .model import MyModel, RawProblem, read_sdpa, raw_from_sdpa_arrays, prepare_model  # noqa: F401
from .solver import (DEFAULT_OPTIONS, MySolver, Halpha, PosDefException, load, solve, Optimizer)  # noqa: F401
from . import problems  # noqa: F401
from . import dist  # noqa: F401
