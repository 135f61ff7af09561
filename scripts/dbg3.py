import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import loraine_oracle as lo, sdpa_io
from loraine_jl_b200 import solver as S
import test_gpu_solver as T
arrays = pkg.problems.multiblock_lp(3, 12, 10, 7)
o = dict(kit=1, preconditioner=1, erank=1, aamat=2, initpoint=1, verb=0, eDIMACS=1e-12, tol_cg=1e-9, tol_cg_min=1e-9)
opt, ora = T.make_pair(pkg, arrays, o)
lo_, s, ha = ora
gs = opt.solver
S.setup_solver(gs, opt.halpha); S.initial_point(gs); lo_.setup_solver(s, ha); lo_.initial_point(s)
for it in range(4):
    for mod, st, hh in ((S, gs, opt.halpha), (lo_, s, ha)):
        st.iter += 1; st.cg_iter_pre = st.cg_iter_cor = 0
        mod.find_mu(st); mod.prepare_W(st)
    S.predictor(gs, opt.halpha); lo_.predictor(s, ha)
    print(it, "pred: cg", gs.cg_iter_pre, s.cg_iter_pre, "exit", gs.cg_exit_code, "dely rel", T.relerr(gs.get_array("DELY"), s.dely), "rhs rel", T.relerr(gs.get_array("RHS"), s.dely*0+1) if False else "", "alpha", gs.alpha, s.alpha, gs.alpha_lin, s.alpha_lin, gs.beta_lin, s.beta_lin)
    sg, so = S.sigma_update(gs), lo_.sigma_update(s)
    S.corrector(gs, opt.halpha); lo_.corrector(s, ha)
    print(it, "corr: cg", gs.cg_iter_cor, s.cg_iter_cor, "exit", gs.cg_exit_code, "dely rel", T.relerr(gs.get_array("DELY"), s.dely), "sigma", sg, so)
    S.check_convergence(gs); lo_.check_convergence(s)
    print(it, "dimacs", gs.DIMACS_error, s.DIMACS_error, [gs.err1,gs.err2,gs.err3,gs.err4,gs.err5,gs.err6], [s.err1,s.err2,s.err3,s.err4,s.err5,s.err6])
