#!/usr/bin/env python
"""Benchmark of the Loraine.jl per-iteration interior-point hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C3|C4|C5|...-mini]

A "step" is ONE interior-point iteration (find_mu, prepare_W, predictor, sigma_update, corrector, check_convergence) of
the workload BASELINE.json's metric is quoted on: configs[1], the synthetic Max-Cut SDP n = 5000 with the rank-one Schur
path (datarank = -1, kit = 0).  `value` = seconds per IP iteration with the problem and the iterate resident in HBM;
`e2e` = the same iteration driven through the C ABI with the iterate in HOST memory (upload X, S, y / download y, X, S
inside the timed region).  Multi-GPU (--gpus N under torchrun): this workload has a single PSD block, the path does not
shard ("replicas only"): every rank runs an independent replica and `value` is job seconds per iteration.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sec/IP-iteration"
UNIT = "s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--cpu-sample-n", type=int, default=2000, help="side of the reduced instance timed on the host CPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the end-to-end solve to convergence reported under `solve`")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([t.strip() for t in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) >= 7 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def algorithmic_flops(md, datarank):
    """SURVEY 8(d).  Rank-one assembly: sum_i n_var^2 m_i (lower SYRK, 2 flop/MAC) + 2 nnz(B_i) m_i + n_var^2/2.
    General assembly: sum_i sum_j min(F1_j, F3_j), F1_j = 2 m nnz_j + 2 m^3 + 2 sum_{k>=j} nnz_k, F3_j = 4 nnz_j sum_{k>=j} nnz_k
    (k over the matrices present in block i in sigmaA order).  Cholesky: n_var^3 / 3."""
    n = float(md.n)
    asm = 0.0
    for i, m in enumerate(md.msizes):
        if datarank == -1 and md.B:
            asm += n * n * m + 2.0 * md.B[i].nnz * m + 0.5 * n * n
        else:
            nz = np.sort(md.nzA[:, i][md.nzA[:, i] > 0])[::-1].astype(np.float64)
            suf = np.cumsum(nz[::-1])[::-1]
            f1 = 2.0 * m * nz + 2.0 * float(m) ** 3 + 2.0 * suf
            f3 = 4.0 * nz * suf
            asm += float(np.minimum(f1, f3).sum())
    if md.nlin:
        asm += 2.0 * float((np.diff(md.C_lin.tocsc().indptr).astype(np.float64) ** 2).sum())
    return dict(assemble=asm, factor=n ** 3 / 3.0)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core.  Returns the BLAS thread count."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=n)
        got = [int(p.get("num_threads", 0)) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(got) if got else n
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", n))


# ----------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path on the host cores.  Julia is not installed in this image, so the
    arm executes the NumPy/LAPACK oracle (oracle/loraine_oracle.py, kind = "port") with all host threads."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import __graft_entry__ as g
    pkg = g.load_package()
    from oracle import loraine_oracle as lo, sdpa_io
    blas_threads = use_all_host_threads()
    cfg = pkg.problems.CONFIGS[args.workload]
    full = cfg["gen"]
    ns = args.cpu_sample_n
    sample_note = ""
    scale = 1.0
    if args.workload == "C2":
        rows = 25
        cols = max(4, ns // rows)
        arrays = pkg.problems.maxcut_torus(rows, cols, 5000)
        nfull = 5000
        scale = (nfull / float(rows * cols)) ** 3
        sample_note = (f"max-cut torus {rows}x{cols} (n = m = {rows * cols}) from the same generator/seed; every phase of the "
                       f"iteration is O(n^3) dense work, seconds scaled by (5000/{rows * cols})^3 = {scale:.1f} to the full size "
                       f"(explicit extrapolation; the full-size CPU iteration takes minutes)")
    else:
        arrays = full()
        sample_note = "full-size instance"
    o = dict(cfg["options"], verb=0)
    md = lo.prepare_model(sdpa_io.raw_from_sdpa_arrays(*arrays), datarank=int(o.get("datarank", 0)), kappa=int(o.get("datasparsity", 8)))
    s, ha = lo.load(md, o)
    lo.setup_solver(s, ha)
    lo.initial_point(s)

    def step():
        lo.myIPstep(s, ha)
        s.tol_cg = max(s.tol_cg * s.tol_cg_up, s.tol_cg_min)
        lo.check_convergence(s)
        if s.status != 0:
            lo.initial_point(s)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    raw = (time.perf_counter() - t0) / args.steps
    cores = blas_threads
    val = raw * scale
    line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=val * 1e3,
                higher_is_better=False, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.workload), sample=sample_note),
                cpu_baseline=dict(value=val, unit=UNIT, cores=cores, kind="port", sample=sample_note, measured_s_per_iteration_on_sample=raw),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def workload_name(w):
    return {"C5": "configs[4]: synthetic large-Schur SDP n_var=40000 constraints, 1 PSD block m=1000, kit=0, datarank=0 (seed 40000)",
            "C4": "configs[3]: synthetic multi-block SDP, 50 PSD blocks of size 200 + LP block of 2000 rows, n_var=10000, kit=0",
            "C3": "configs[2]: synthetic Lovasz-theta SDP (thetaG11 layout) m=801, n_var=2401, kit=1 CG, preconditioner=1",
            "C2": "configs[1]: synthetic Max-Cut SDP (ex_maxcut.jl / maxG11 layout) n=5000, 1 PSD block m=5000, datarank=-1 rank-one "
                  "Schur path, kit=0 (torus 50x100, +-1 weights, seed 5000)"}.get(w, w)


# ----------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    rank, world, local = dist_env()
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout for the one JSON line: NCCL's own messages (version banner, INFO) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    pkg = g.load_package()
    from loraine_jl_b200 import solver as S, _lib
    L = _lib.lib()
    L.lrn_dbg_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.lrn_dbg_gemm_profile.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]

    cfg = pkg.problems.CONFIGS[args.workload]
    arrays = cfg["gen"]()
    opt = pkg.Optimizer()
    for k, v in dict(cfg["options"], verb=0, device=local).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    s, ha = opt.solver, opt.halpha
    S.setup_solver(s, ha)
    sharded = False
    if world > 1 and args.workload.startswith("C5"):
        sharded = pkg.dist.init_distributed(s)      # Schur assembly + Cholesky sharded over the ranks (NCCL panel broadcast)
    S.initial_point(s)
    md = s.model

    def step():
        S.myIPstep(s, ha)
        s.itertime = 0.0
        s.tol_cg = max(s.tol_cg * s.tol_cg_up, s.tol_cg_min)
        S.check_convergence(s)
        if s.status != 0:               # converged (or failed): restart the same solve so that every step is a real iteration
            S.initial_point(s)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tol_cg0 = s.tol_cg
    for _ in range(max(args.warmup, 3)):
        step()
    # ---- timed region: device-resident iterations -------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    s.timers(reset=True)
    launches0 = L.lrn_kernel_launches()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    t_dev = time.perf_counter() - t0
    launches = L.lrn_kernel_launches() - launches0
    phase = s.timers(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    # ---- e2e: the iterate lives in HOST memory; upload before / download after every iteration ------------------------
    # same iterations as the device-timed region: restart from the initial point, run the same warm-up, time the same K steps
    S.initial_point(s)
    s.tol_cg = tol_cg0
    y, X, xl = S.get_solution(s)
    PD = C.POINTER(C.c_double)
    Sm = [np.zeros((m, m), order="F") for m in md.msizes]
    sl = np.zeros(md.nlin)
    Sp = (PD * max(1, md.nlmi))(*[x.ctypes.data_as(PD) for x in Sm])
    s._call("lrn_get_slack", Sp, sl.ctypes.data_as(PD) if md.nlin else None)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    Xh = [np.asfortranarray(pin(x.T).T) for x in X]
    Sh = [np.asfortranarray(pin(x.T).T) for x in Sm]
    yh, xlh, slh = pin(y), pin(xl), pin(sl)
    Shp = (PD * max(1, md.nlmi))(*[x.ctypes.data_as(PD) for x in Sh])
    h2d = sum(x.nbytes for x in Xh) + sum(x.nbytes for x in Sh) + yh.nbytes + xlh.nbytes + slh.nbytes
    d2h = h2d
    e2e_steps = args.steps

    def e2e_step():
        S.set_iterate(s, Xh, Sh, yh, xlh, slh)                      # H2D from the pinned buffers
        S.myIPstep(s, ha)
        s.itertime = 0.0
        s.tol_cg = max(s.tol_cg * s.tol_cg_up, s.tol_cg_min)
        S.check_convergence(s)
        if s.status != 0:
            S.initial_point(s)
        S.get_solution(s, out=(yh, Xh, xlh))                        # D2H straight into the pinned buffers
        s._call("lrn_get_slack", Shp, slh.ctypes.data_as(PD) if md.nlin else None)

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    # ---- roofline pass: per-launch CUDA events around the dominant kernel (the DMMA GEMM) ---------------------------
    ms, fl, nl = C.c_double(), C.c_double(), C.c_int64()
    L.lrn_dbg_gemm_profile(1, None, None, None)
    prof_steps = 1
    for _ in range(prof_steps):
        step()
    L.lrn_dbg_gemm_profile(0, C.byref(ms), C.byref(fl), C.byref(nl))
    peak = C.c_double()
    L.lrn_dbg_peak(0, C.byref(peak))
    # ---- end-to-end solve to the reference's stopping rule (outside the timed regions; every rank runs it) ---------------
    solve_info = None
    if not args.no_solve:
        barrier()
        t0 = time.perf_counter()
        S.solve(s, ha)
        barrier()
        solve_info = dict(seconds=time.perf_counter() - t0, iterations=int(s.iter), status=int(s.status),
                          dimacs_error=float(s.DIMACS_error), primal_obj=float(s.primal_obj), dual_obj=float(s.dual_obj),
                          cg_iterations=int(s.cg_iter_tot) if s.kit == 1 else None)

    tmax = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(tmax[0]), float(tmax[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # replicas: the job advances `world` independent solves per step; sharded: all ranks advance ONE solve together
    units = 1 if sharded else world
    sec_per_iter = t_dev / (args.steps * units)
    e2e_val = t_e2e / (e2e_steps * units)
    alg = algorithmic_flops(md, int(s.datarank))
    asm_ms = phase["schur_assemble"][0] / max(1, phase["schur_assemble"][1])
    fac_ms = phase["schur_factor"][0] / max(1, phase["schur_factor"][1])
    achieved = fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0.0
    # per-kernel split of the same profile: the roofline object describes the kernel with the largest share of the step
    fams = []
    for code, kname in ((10, "dgemm_dmma_kernel (cp.async-fed DMMA GEMM: congruences, Gram products, TRSM, edge strips)"),
                        (11, "dgemm_dmma_bulk_kernel (TMA-fed DMMA GEMM: Schur SYRK, Cholesky trailing updates, large congruences)"),
                        (12, "panel_rotate_kernel (TMA-fed DMMA panel rotation of the block-Jacobi SVD)")):
        fm, ff, fn = C.c_double(), C.c_double(), C.c_int64()
        L.lrn_dbg_gemm_profile(code, C.byref(fm), C.byref(ff), C.byref(fn))
        if fn.value > 0 and fm.value > 0:
            fams.append(dict(kernel=kname, launches=int(fn.value), kernel_ms_per_step=fm.value / prof_steps,
                             algorithmic_flops_per_launch=ff.value / fn.value,
                             avg_launch_us=1e3 * fm.value / fn.value, achieved=ff.value / (fm.value * 1e-3) / 1e12))
    dom = max(fams, key=lambda d: d["kernel_ms_per_step"]) if fams else None
    line = dict(
        metric=METRIC, value=sec_per_iter, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=1e3 * t_dev / args.steps, higher_is_better=False, scaling="strong" if sharded else "weak", vs_baseline=None,
        dtype="f64", data="synthetic",
        config=dict(workload=workload_name(args.workload), n_var=md.n, msizes=md.msizes[:4], nlin=md.nlin,
                    options=cfg["options"], l2="inputs larger than L2 (every dense operand is 200 MB; 21 resident m x m matrices)",
                    parallelism=("schur assembly + Cholesky sharded block-cyclic over %d GPUs (NCCL panel broadcast), rest replicated" % world)
                    if sharded else ("replicas only" if world > 1 else "single GPU")),
        e2e=dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), steps=e2e_steps),
        gpu_launches=int(launches),
        clocks=clocks,
        solve=solve_info,
        phases_ms_per_iteration={k: v[0] / args.steps for k, v in phase.items()},
        # CUDA-event time of the ABI calls of one step on the library stream (nested phases svd / eigmin not double counted);
        # `value` is the host clock around the same steps between device synchronisations and also contains the host logic
        device_event_ms_per_step=sum(v[0] for k, v in phase.items() if k not in ("svd", "eigmin")) / args.steps,
        schur=dict(assemble_ms=asm_ms, assemble_tflops=alg["assemble"] / (asm_ms * 1e-3) / 1e12 if asm_ms > 0 else None,
                   factor_ms=fac_ms, factor_tflops=alg["factor"] / (fac_ms * 1e-3) / 1e12 if fac_ms > 0 else None,
                   assemble_plus_factor_tflops=(alg["assemble"] + alg["factor"]) / ((asm_ms + fac_ms) * 1e-3) / 1e12
                   if asm_ms + fac_ms > 0 else None),
        roofline=dict(bound="tensor", kernel=dom["kernel"] if dom else None,
                      achieved=dom["achieved"] if dom else None, peak=peak.value, unit="TFLOP/s",
                      frac=dom["achieved"] / peak.value if dom and peak.value else None,
                      # DRAM bytes per launch from the ncu --set full capture of the same shape (profiles/r1c_svd_round_ncu_full.csv:
                      # 212.7 MB read + 161.1 MB written; algorithmic: 200 MB read + 200 MB written incl. padding rows)
                      traffic=373.8e6 if dom and dom["kernel"].startswith("panel_rotate") and args.workload == "C2" else None,
                      launches_profiled=dom["launches"] if dom else 0, kernel_ms_per_step=dom["kernel_ms_per_step"] if dom else None,
                      avg_launch_us=dom["avg_launch_us"] if dom else None,
                      algorithmic_flops_per_launch=dom["algorithmic_flops_per_launch"] if dom else None,
                      all_dmma_kernels=dict(achieved=achieved, launches=int(nl.value), kernel_ms_per_step=ms.value / prof_steps),
                      by_kernel=fams,
                      peak_source="measured in this run: register-resident mma.sync.m8n8k4.f64 loop on all SMs "
                                  "(MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM 8192^3 on this pool: 35.4 TFLOP/s)"),
        stats=s.stats(),
    )
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(pkg, args)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(pkg, args):
    """oracle (kind "port") timed on this box's host cores on a bounded sample of the same workload"""
    from oracle import loraine_oracle as lo, sdpa_io
    blas_threads = use_all_host_threads()
    cfg = pkg.problems.CONFIGS[args.workload]
    scale, note = 1.0, "full-size instance"
    if args.workload == "C2":
        rows, cols = 25, max(4, args.cpu_sample_n // 25)
        arrays = pkg.problems.maxcut_torus(rows, cols, 5000)
        scale = (5000.0 / (rows * cols)) ** 3
        note = (f"1 IP iteration (after 1 warm-up iteration) of the same generator at torus {rows}x{cols} (n = m = {rows * cols}); all phases are "
                f"O(n^3): seconds scaled by (5000/{rows * cols})^3 = {scale:.0f} to the full size (explicit extrapolation)")
    elif args.workload == "C5":
        ns = 4000
        arrays = pkg.problems.large_schur(1000, ns, 40000)
        scale = (40000.0 / ns) ** 3
        note = (f"1 IP iteration (after 1 warm-up iteration) of the same generator at reduced n_var = {ns} (m = 1000 kept); the "
                f"iteration is dominated by the n_var^3/3 Cholesky and the O(n_var^2) pair assembly: seconds scaled by "
                f"(40000/{ns})^3 = {scale:.0f} (explicit extrapolation, upper estimate)")
    else:
        arrays = cfg["gen"]()
    o = dict(cfg["options"], verb=0)
    md = lo.prepare_model(sdpa_io.raw_from_sdpa_arrays(*arrays), datarank=int(o.get("datarank", 0)), kappa=int(o.get("datasparsity", 8)))
    s, ha = lo.load(md, o)
    lo.setup_solver(s, ha)
    lo.initial_point(s)
    lo.myIPstep(s, ha)
    lo.check_convergence(s)
    t0 = time.perf_counter()
    lo.myIPstep(s, ha)
    lo.check_convergence(s)
    raw = time.perf_counter() - t0
    return dict(value=raw * scale, unit=UNIT, cores=blas_threads, kind="port", sample=note, measured_s_per_iteration_on_sample=raw)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
