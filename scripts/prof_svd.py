"""Time one block-Jacobi SVD of an IPM-like matrix (cond ~ 4) through the debug hook; used under ncu for the per-round split."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
L = pkg._lib.lib()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rng = np.random.default_rng(1)
Q1, _ = np.linalg.qr(rng.standard_normal((m, m)))
Q2, _ = np.linalg.qr(rng.standard_normal((m, m)))
s = np.linspace(1.0, 4.0, m)
A = np.asfortranarray((Q1 * s) @ Q2.T)
UD = np.zeros((m, m), order="F"); sg = np.zeros(m); sw = C.c_int32(0); ms = C.c_double(0)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
L.lrn_dbg_svd.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                          C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    rc = L.lrn_dbg_svd(m, dp(A), dp(UD), None, dp(sg), float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8, C.byref(sw), C.byref(ms))
    print("svd m=%d rc=%d %.1f ms sweeps %d (%.2f ms/sweep) maxrel %.2e" % (m, rc, ms.value, sw.value, ms.value / max(sw.value, 1),
          np.abs(np.sort(sg)[::-1] - s[::-1]).max()), flush=True)
