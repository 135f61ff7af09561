"""Factor + solve of an n x n diagonally dominant SPD matrix through the debug hook (ncu target for the triangular-solve kernels):
   ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:trsv python scripts/prof_trsv.py 20000"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from loraine_jl_b200 import _lib  # noqa: E402

L = _lib.lib()
pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
L.lrn_dbg_cholesky.argtypes = [C.c_int32, pd, pd, C.c_int32, pi, C.c_int32, pd]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(0)
A = rng.random((n, n))
A = np.asfortranarray(A + A.T)
A[np.diag_indices(n)] += 2.0 * n
x = rng.standard_normal(n)
b = x.copy()
info = C.c_int32()
L.lrn_dbg_cholesky(n, A.ctypes.data_as(pd), x.ctypes.data_as(pd), 3, C.byref(info), 0, None)
print("info", info.value)
