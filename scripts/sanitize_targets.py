"""Small invocations of the kernels that synchronise through mbarriers / bulk copies / grid barriers / shared-memory staging,
meant to run under compute-sanitizer:
    compute-sanitizer --tool racecheck python scripts/sanitize_targets.py
    compute-sanitizer --tool memcheck  python scripts/sanitize_targets.py
Targets: dgemm_dmma_bulk_kernel (A B^T and A B), panel_factor_kernel + look-ahead Cholesky, panel_rotate_kernel and
jacobi_cross64_reg_kernel (block-Jacobi SVD, two sweeps), k_schur_pairs_staged (sparse Schur assembly), the triangular solves."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from loraine_jl_b200 import _lib, solver as S  # noqa: E402

L = _lib.lib()
i32, dbl = C.c_int32, C.c_double
pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
L.lrn_dbg_gemm.argtypes = [i32, i32, i32, i32, i32, dbl, pd, pd, dbl, pd, i32, i32, pd, i32, i32, pd]
L.lrn_dbg_cholesky.argtypes = [i32, pd, pd, i32, pi, i32, pd]
L.lrn_dbg_svd.argtypes = [i32, pd, pd, pd, pd, dbl, pi, pd]
dp = lambda a: a.ctypes.data_as(pd)
rng = np.random.default_rng(0)
for (M, N, K, tb) in ((1536, 1408, 256, 1), (1536, 1408, 256, 0)):
    A = np.asfortranarray(rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((N, K) if tb else (K, N)))
    Cm = np.asfortranarray(np.zeros((M, N)))
    assert L.lrn_dbg_gemm(M, N, K, 0, tb, 1.0, dp(A), dp(B), 0.0, dp(Cm), 0, 0, None, 0, 0, None) == 0
    ref = A @ (B.T if tb else B)
    print("gemm", M, N, K, tb, "relerr %.2e" % (np.linalg.norm(Cm - ref) / np.linalg.norm(ref)), flush=True)
n = 1300
Gm = rng.standard_normal((n, n))
A = np.asfortranarray(Gm @ Gm.T / n + np.eye(n))
A0 = A.copy()
x = rng.standard_normal(n)
b = x.copy()
info = C.c_int32()
assert L.lrn_dbg_cholesky(n, dp(A), dp(x), 3, C.byref(info), 0, None) == 0 and info.value == 0
print("chol", n, "solve relerr %.2e" % (np.linalg.norm(A0 @ x - b) / np.linalg.norm(b)), flush=True)
m = 1100
A = np.asfortranarray(rng.standard_normal((m, m)))
UD, sg = np.asfortranarray(np.zeros((m, m))), np.zeros(m)
sw, ms = C.c_int32(), C.c_double()
assert L.lrn_dbg_svd(m, dp(A), dp(UD), None, dp(sg), 0.3, C.byref(sw), C.byref(ms)) == 0     # loose tolerance: two sweeps
print("svd", m, "sweeps", sw.value, flush=True)
arrays = pkg.problems.large_schur(60, 1500, 40000)
opt = pkg.Optimizer()
for k, v in dict(kit=0, datarank=0, initpoint=1, verb=0).items():
    opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
s = opt.solver
S.setup_solver(s, opt.halpha)
S.initial_point(s)
s.iter = 1
S.find_mu(s); S.prepare_W(s); s._call("lrn_residuals")
for mode in (1.0, 0.0):
    s._call("lrn_set_option", b"pair_kernel", mode)
    s._call("lrn_schur_assemble")
s._call("lrn_rhs_predictor")
assert s._call("lrn_schur_factor") == 0
s._call("lrn_schur_solve", 3)
print("assemble + factor + solve ok", flush=True)
s.close()
