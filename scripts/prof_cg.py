"""Per-IP-iteration CG cost on C3 (theta, m = 801, H_alpha): milliseconds in lrn_pcg and CG iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
S = pkg.solver
cfg = pkg.problems.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
opt = pkg.Optimizer()
for k, v in dict(cfg["options"], verb=0, device=0).items():
    opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(*cfg["gen"]()))
s, ha = opt.solver, opt.halpha
S.setup_solver(s, ha)
S.initial_point(s)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    s.timers(reset=True)
    c0 = s.cg_iter_tot
    S.myIPstep(s, ha)
    s.itertime = 0.0
    s.tol_cg = max(s.tol_cg * s.tol_cg_up, s.tol_cg_min)
    S.check_convergence(s)
    t = s.timers(reset=True)
    n = s.cg_iter_tot - c0
    print("it %2d  cg %7.2f ms  %4d CG iterations  %.3f ms/CG-iteration   prec_prepare %.2f ms  prepare_W %.2f ms" % (
        s.iter, t["cg"][0], n, t["cg"][0] / max(n, 1), t["prec_prepare"][0], t["prepare_W"][0]), flush=True)
    if s.status != 0:
        break

# ---- operator / preconditioner apply in isolation (includes one H2D + one D2H of n_var doubles per call) ----
import ctypes as C, time, numpy as np, torch
x = np.random.default_rng(0).standard_normal(s.model.n); out = np.zeros_like(x)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
s._call("lrn_prec_prepare", 1)
for kind, name in ((-1, "operator A"), (1, "preconditioner H_alpha"), (0, "identity (copy only: call overhead)")):
    for _ in range(5):
        s._call("lrn_apply_operator", kind, dp(x), dp(out))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        s._call("lrn_apply_operator", kind, dp(x), dp(out))
    torch.cuda.synchronize()
    print("%-40s %.1f us per call" % (name, (time.perf_counter() - t0) / 200 * 1e6), flush=True)
