"""Double-double LP path: device time per phase on a synthetic LP (default n = 2000 multipliers, 5000 LP rows) and the rate of
the double-double Cholesky in dd multiply-adds per second (n^3/6 of them).
usage: python scripts/prof_ddlp.py [n] [nlin] [density] [iterations]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as g
from dd_common import random_lp

pkg = g.load_package()
from loraine_jl_b200 import dd_lp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nlin = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
dens = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
spec = random_lp(n, nlin, 1, density=dens)
md = pkg.prepare_model(pkg.RawProblem(**spec))
s = dd_lp.DDSolver(md, dict(pkg.DEFAULT_OPTIONS, eDIMACS=1e-24, verb=0))
dd_lp.setup_solver(s)
dd_lp.initial_point(s)
for it in range(iters):
    s.timers(reset=True)
    t0 = time.perf_counter()
    dd_lp.myIPstep(s)
    s.itertime = time.perf_counter() - t0
    dd_lp.check_convergence(s)
    t = s.timers()
    fma = n ** 3 / 6
    print("it %d: %.1f ms wall | assemble %.2f  factor %.2f (%.1f G dd-FMA/s)  solve %.2f  other %.2f ms | DIMACS %.3e" % (
        s.iter, s.itertime * 1e3, t["schur_assemble"], t["schur_factor"], fma / (t["schur_factor"] * 1e-3) / 1e9, t["schur_solve"],
        t["other"], float(s.DIMACS_error)), flush=True)
t0 = time.perf_counter()
dd_lp.solve(s, setup=False)
print("solve: status %d after %d iterations, DIMACS %.3e, %.2f s" % (s.status, s.iter, float(s.DIMACS_error), time.perf_counter() - t0))
