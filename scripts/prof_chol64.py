import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from loraine_jl_b200 import _lib
L = _lib.lib()
i32, dbl = C.c_int32, C.c_double; pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
L.lrn_dbg_cholesky.argtypes = [i32, pd, pd, i32, pi, i32, pd]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
G = rng.standard_normal((n, n)); A = np.asfortranarray(G @ G.T / n + np.eye(n))
info, ms = C.c_int32(), C.c_double()
L.lrn_dbg_cholesky(n, A.ctypes.data_as(pd), None, 0, C.byref(info), 3, C.byref(ms))
print("chol", n, ms.value, "ms info", info.value)
