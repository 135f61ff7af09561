"""Runs the named hot kernels once each in isolation (for `ncu --set full` captures):
  1. rank-one Schur assembly: SpMM-free SYRK with squared epilogue  H += ((BG)(BG)').^2, n_var = m = 5000 (C2 shape)
  2. Cholesky trailing update shape  C -= P P'  (lower, K = 256)
  3. block-Jacobi panel update shape  (M = 10056, N = K = 64, batch 79)  -- plain NN GEMM with the same tile config
"""
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
i32, dbl = C.c_int32, C.c_double
pd = C.POINTER(C.c_double)
L.lrn_dbg_gemm.argtypes = [i32, i32, i32, i32, i32, dbl, pd, pd, dbl, pd, i32, i32, pd, i32, i32, pd]
dp = lambda a: a.ctypes.data_as(pd)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
Kp = (n + 31) // 32 * 32          # the solver zero-pads B*G to K % 32 == 0: TMA-fed kernel + edge strips
A = np.asfortranarray(np.zeros((n, Kp)))
A[:, :n] = rng.standard_normal((n, n))
H = np.asfortranarray(np.zeros((n, n)))
ms = C.c_double()
L.lrn_dbg_gemm(n, n, Kp, 0, 1, 1.0, dp(A), dp(A), 1.0, dp(H), 1, 1, None, 0, 3, C.byref(ms))
print("syrk-square lower %d: %.3f ms  %.2f TFLOP/s (algorithmic n^3)" % (n, ms.value, n ** 3 / (ms.value * 1e-3) / 1e12))
P = np.asfortranarray(rng.standard_normal((n, 256)))
L.lrn_dbg_gemm(n, n, 256, 0, 1, -1.0, dp(P), dp(P), 1.0, dp(H), 0, 1, None, 0, 3, C.byref(ms))
print("trailing update lower %d K=256: %.3f ms  %.2f TFLOP/s" % (n, ms.value, n * n * 256 / (ms.value * 1e-3) / 1e12))
