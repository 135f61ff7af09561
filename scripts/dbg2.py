import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from oracle import loraine_oracle as lo, sdpa_io
arrays = pkg.problems.theta_torus(6, 8)
o = dict(pkg.problems.CONFIGS["C3-mini"]["options"], verb=2, preconditioner=4)
ref = lo.solve_raw(sdpa_io.raw_from_sdpa_arrays(*arrays), dict(o, verb=1))
print("oracle", ref.iter, ref.primal_obj, [ (t["cg_pre"], t["cg_cor"]) for t in ref.trace])
opt = pkg.Optimizer()
for k, v in o.items(): opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
opt.optimize()
print(opt.solver.status, opt.solver.stats(), [(t["cg_pre"], t["cg_cor"]) for t in opt.solver.trace])
# sweeps on a realistic problem
z = np.load(os.path.join(ROOT, "tests/golden/maxG11.npz"))
opt = pkg.Optimizer()
for k, v in dict(kit=0, datarank=-1, initpoint=1, eDIMACS=1e-6, verb=1).items(): opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(int(z["n"]), [int(b) for b in z["bs"]], z["c"], z["body"]))
from loraine_jl_b200 import solver as S
s = opt.solver
S.setup_solver(s, opt.halpha); S.initial_point(s)
import time
for it in range(16):
    t = time.time(); S.myIPstep(s, opt.halpha); s.itertime = time.time() - t
    S.check_convergence(s)
    print("  stats", s.stats(), {k: round(v[0], 1) for k, v in s.timers(reset=True).items()})
    if s.status: break
print("maxG11", s.iter, s.primal_obj, "oracle", int(z["oracle_iters"]), float(z["oracle_obj"]))
