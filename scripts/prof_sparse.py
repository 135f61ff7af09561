"""Runs the HBM/L2-bound kernels of the path once on a C5-shaped instance (sparse-pair Schur assembly, A / A' operators,
Lanczos GEMV) for `ncu --set full` captures."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from loraine_jl_b200 import solver as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
arrays = pkg.problems.large_schur(1000, n, 40000)
opt = pkg.Optimizer()
for k, v in dict(kit=0, datarank=0, initpoint=1, eDIMACS=1e-6, verb=0).items():
    opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
s = opt.solver
S.setup_solver(s, opt.halpha); S.initial_point(s)
for it in range(2):
    S.myIPstep(s, opt.halpha); S.check_convergence(s)
t = s.timers(reset=True)
S.find_mu(s); S.prepare_W(s); s._call("lrn_residuals"); s._call("lrn_schur_assemble"); s._call("lrn_rhs_predictor")
t = s.timers()
print("assemble ms", t["schur_assemble"], "residuals", t["residuals"], "rhs", t["rhs"])
