"""ncu target for the kit = 1 path on C3: ONE Schur operator apply (MyA) and ONE H_alpha preconditioner apply (MyM) between
cudaProfilerStart/Stop, after two IP iterations (`ncu --profile-from-start off --set full ... python scripts/prof_cg_ops.py`)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

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
for _ in range(2):
    S.myIPstep(s, ha)
    S.check_convergence(s)
x = np.random.default_rng(0).standard_normal(s.model.n)
out = np.zeros_like(x)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
s._call("lrn_prec_prepare", 1)
for kind in (-1, 1):
    for _ in range(3):
        s._call("lrn_apply_operator", kind, dp(x), dp(out))
torch.cuda.synchronize()
torch.cuda.profiler.start()
s._call("lrn_apply_operator", -1, dp(x), dp(out))
s._call("lrn_apply_operator", 1, dp(x), dp(out))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("n_var", s.model.n, "done")
