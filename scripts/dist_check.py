"""torchrun --nproc-per-node N scripts/dist_check.py [workload]: solves the workload with the Schur assembly and the Cholesky
sharded over N GPUs and compares against the single-GPU run on rank 0 (H parity, dely parity, iteration counts)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as g

pkg = g.load_package()
from loraine_jl_b200 import solver as S

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
name = sys.argv[1] if len(sys.argv) > 1 else "C5-mini"
cfg = pkg.problems.CONFIGS[name]
arrays = cfg["gen"]()


def make(distributed):
    opt = pkg.Optimizer()
    for k, v in dict(cfg["options"], verb=0, device=local).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    s = opt.solver
    S.setup_solver(s, opt.halpha)
    if distributed:
        assert pkg.dist.init_distributed(s)
    S.initial_point(s)
    return opt, s


opt_d, sd = make(True)
opt_1, s1 = make(False)
for it in range(3):
    for s, o in ((sd, opt_d), (s1, opt_1)):
        s.iter += 1; s.cg_iter_pre = s.cg_iter_cor = 0
        S.find_mu(s); S.prepare_W(s)
        s._call("lrn_residuals"); s._call("lrn_schur_assemble")
    Hd = sd.get_array("H")            # all-reduced over ranks inside the library
    H1 = s1.get_array("H")
    eH = np.linalg.norm(Hd - H1) / np.linalg.norm(H1)
    for s, o in ((sd, opt_d), (s1, opt_1)):
        s._call("lrn_rhs_predictor"); s._call("lrn_schur_factor")
        s.cholBBBB = S._DeviceFactor(s, False); s.cholBBBB.solve_reference_expression()
    dd, d1 = sd.get_array("DELY"), s1.get_array("DELY")
    eD = np.linalg.norm(dd - d1) / np.linalg.norm(d1)
    Ld, L1 = sd.get_array("L"), s1.get_array("L")
    eL = np.linalg.norm(Ld - L1) / np.linalg.norm(L1)
    for s, o in ((sd, opt_d), (s1, opt_1)):
        s.predict = True
        S.find_step(s); S.sigma_update(s); S.corrector(s, o.halpha); S.check_convergence(s)
    print(f"[rank {rank}] it {it + 1}: H relerr {eH:.2e}  L relerr {eL:.2e}  dely relerr {eD:.2e}  dimacs {sd.DIMACS_error:.3e} vs {s1.DIMACS_error:.3e}", flush=True)
    assert eH < 1e-12 and eL < 1e-10 and eD < 1e-8
dist.barrier()
if rank == 0:
    print("dist_check ok", name, "world", world)
dist.destroy_process_group()
