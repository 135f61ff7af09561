"""Cholesky of one n x n SPD matrix through the debug hook (ncu target: `ncu --metrics gpu__time_duration.sum ... python scripts/prof_chol.py 5000`)."""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
reps = max(1, int(sys.argv[2])) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
Gm = rng.standard_normal((n, n))
A = np.asfortranarray(Gm @ Gm.T / n + np.eye(n))
info, ms = C.c_int32(), C.c_double()
L.lrn_dbg_cholesky(n, A.ctypes.data_as(pd), None, 0, C.byref(info), reps, C.byref(ms))
print("chol", n, "%.3f ms %.2f TF/s info %d" % (ms.value, n ** 3 / 3 / (ms.value * 1e-3) / 1e12, info.value))
