import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from loraine_jl_b200 import _lib
L = _lib.lib()
i32, dbl = C.c_int32, C.c_double; pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
L.lrn_dbg_svd.argtypes = [i32, pd, pd, pd, pd, dbl, pi, pd]
dp = lambda a: a.ctypes.data_as(pd)
for m in (65, 130, 200):
    rng = np.random.default_rng(m)
    A = np.asfortranarray(rng.standard_normal((m, m)))
    UD, V, sg = np.asfortranarray(np.zeros((m, m))), np.asfortranarray(np.zeros((m, m))), np.zeros(m)
    sw, ms = C.c_int32(), C.c_double()
    L.lrn_dbg_svd(m, dp(A), dp(UD), dp(V), dp(sg), 0.0, C.byref(sw), C.byref(ms))
    ref = np.linalg.svd(A, compute_uv=False)
    print(m, "sweeps", sw.value, "maxrel", np.max(np.abs(sg-ref)/ref), "orthV", np.linalg.norm(V.T@V-np.eye(m)), "AV-UD", np.linalg.norm(A@V-UD), "sg[:5]", sg[:5], ref[:5])
cfg = pkg.problems.CONFIGS["C2-mini"]
opt = pkg.Optimizer()
for k, v in dict(cfg["options"], verb=2, maxit=30).items(): opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(*cfg["gen"]()))
opt.optimize()
print(opt.solver.stats(), opt.solver.status)
