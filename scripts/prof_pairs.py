"""configs[4] Schur assembly alone (ncu target for the staged pair kernel):
   ncu --set full --clock-control none --import-source on -k regex:k_schur_pairs_staged -c 1 -o gpurun_out/pairs python scripts/prof_pairs.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from loraine_jl_b200 import solver as S  # noqa: E402

cfg = pkg.problems.CONFIGS["C5"]
opt = pkg.Optimizer()
for k, v in dict(cfg["options"], verb=0).items():
    opt.set_attribute(k, v)
opt.copy_to(pkg.raw_from_sdpa_arrays(*cfg["gen"]()))
s = opt.solver
S.setup_solver(s, opt.halpha)
S.initial_point(s)
S.myIPstep(s, opt.halpha)
S.check_convergence(s)
s.iter += 1
S.find_mu(s); S.prepare_W(s); s._call("lrn_residuals")
for mode in (1.0, 0.0):
    s._call("lrn_set_option", b"pair_kernel", mode)
    s.timers(reset=True)
    for _ in range(3):
        s._call("lrn_schur_assemble")
    t = s.timers()["schur_assemble"]
    print("pair_kernel", mode, "assemble ms", t[0] / t[1])
