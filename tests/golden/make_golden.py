"""Generates tests/golden/*.npz from the SDPLIB instances bundled with the reference (/root/reference/examples/data).
Run in the build container (the reference tree does not exist on the GPU box):  python tests/golden/make_golden.py

Each file holds the parsed SDPA arrays (n, bs, c, body) -- the public SDPLIB test data, not reference source code -- plus
the oracle's results for that instance (iteration count, objective trace, and for theta1 the Schur matrix / NT scaling /
predictor direction of IP iteration 3).  Julia is not available, so these are ORACLE outputs: they pin regressions and
CUDA-vs-oracle parity; the reference itself pins only the end-to-end optima asserted in tests/test_oracle_golden.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import loraine_oracle as lo, sdpa_io  # noqa: E402

DATA = "/root/reference/examples/data"
OPTS = dict(kit=0, tol_cg=1e-2, tol_cg_min=1e-6, eDIMACS=1e-6, preconditioner=1, erank=1, aamat=2, verb=0, datarank=0,
            initpoint=1, maxit=100, datasparsity=8)

for name in ("theta1", "control1", "tru3", "vib3", "maxG11"):
    n, bs, c, body = sdpa_io.parse_sdpa(os.path.join(DATA, name + ".dat-s"))
    raw = sdpa_io.raw_from_sdpa_arrays(n, bs, c, body)
    opts = dict(OPTS, datarank=-1) if name == "maxG11" else OPTS
    extra = {}
    if name == "theta1":
        md = lo.prepare_model(raw, 0, 8)
        s, ha = lo.load(md, opts)
        s.hooks["H"] = lambda s_, H: extra.__setitem__("H%d" % s_.iter, H.copy())
        s.hooks["W"] = lambda s_: extra.__setitem__("W%d" % s_.iter, s_.W[0].copy())
        s.hooks["dely_pred"] = lambda s_, h, d: extra.__setitem__("dely%d" % s_.iter, d.copy())
        lo.solve(s, ha, max_iters=3)
        extra = {k: v for k, v in extra.items() if k.endswith("3")}
    s = lo.solve_raw(raw, opts)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=n, bs=np.array(bs), c=c, body=body, oracle_iters=s.iter,
                        oracle_obj=s.primal_obj, oracle_dual_obj=s.dual_obj,
                        oracle_obj_trace=np.array([t["obj"] for t in s.trace]),
                        oracle_dimacs_trace=np.array([t["dimacs"] for t in s.trace]), **extra)
    print(name, s.iter, s.primal_obj, s.status)

# Larger instances of the reference's data directory: end-to-end oracle results only (the oracle needs 9 min for tru9,
# 31 min for vib9 and 1.5 min for thetaG11 on 16 cores; thetaG11 is the reference's kit = 1 use case, SDPLIB optimum 400).
for name, extra in (("tru9", {}), ("vib9", {}), ("thetaG11", dict(kit=1))):
    n, bs, c, body = sdpa_io.parse_sdpa(os.path.join(DATA, name + ".dat-s"))
    s = lo.solve_raw(sdpa_io.raw_from_sdpa_arrays(n, bs, c, body), dict(OPTS, **extra))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=n, bs=np.array(bs), c=c, body=body, oracle_iters=s.iter,
                        oracle_obj=s.primal_obj, oracle_dual_obj=s.dual_obj, oracle_cg_iters=getattr(s, "cg_iter_tot", 0))
    print(name, s.iter, s.primal_obj, s.status)

