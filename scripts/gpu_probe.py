"""Device probe: FP64 peaks (DMMA / DFMA register-resident loops, cuBLAS DGEMM as calibrator), HBM copy, and the
library's DMMA GEMM / Cholesky / SVD timings.  Writes gpurun_out/probe.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
from loraine_jl_b200 import _lib  # noqa: E402

L = _lib.lib()
i32, dbl = C.c_int32, C.c_double
pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
L.lrn_dbg_gemm.argtypes = [i32, i32, i32, i32, i32, dbl, pd, pd, dbl, pd, i32, i32, pd, i32, i32, pd]
L.lrn_dbg_cholesky.argtypes = [i32, pd, pd, i32, pi, i32, pd]
L.lrn_dbg_svd.argtypes = [i32, pd, pd, pd, pd, dbl, pi, pd]
L.lrn_dbg_peak.argtypes = [i32, pd]
dp = lambda a: a.ctypes.data_as(pd)
out = {}
v = C.c_double()
for kind, name in ((0, "dmma_tflops"), (1, "dfma_tflops"), (2, "hbm_copy_gbs")):
    L.lrn_dbg_peak(kind, C.byref(v))
    out[name] = v.value
    print(name, v.value, flush=True)
try:
    import torch
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    out["cublas_dgemm_8192_tflops"] = 3 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    print("cublas dgemm 8192", out["cublas_dgemm_8192_tflops"], flush=True)
    del a, b
    torch.cuda.empty_cache()
except Exception as e:  # pragma: no cover
    print("torch probe failed", e)
rng = np.random.default_rng(0)
res = []
sizes = [(4096, 4096, 4096, 0, 0), (4096, 4096, 4096, 0, 1), (4096, 4096, 4096, 1, 0), (8192, 8192, 512, 0, 1), (8192, 8192, 128, 0, 1),
         (5000, 5000, 5000, 0, 1), (10056, 64, 64, 0, 0), (801, 801, 801, 0, 0), (200, 200, 200, 0, 0)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sizes = sizes[:2] + sizes[-3:]
for (M, N, K, ta, tb) in sizes:
    A = np.asfortranarray(rng.standard_normal((K, M) if ta else (M, K)))
    B = np.asfortranarray(rng.standard_normal((N, K) if tb else (K, N)))
    Cm = np.asfortranarray(np.zeros((M, N)))
    ms = C.c_double()
    L.lrn_dbg_gemm(M, N, K, ta, tb, 1.0, dp(A), dp(B), 0.0, dp(Cm), 0, 0, None, 0, 5, C.byref(ms))
    tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12
    res.append(dict(M=M, N=N, K=K, ta=ta, tb=tb, ms=ms.value, tflops=tf))
    print("gemm", M, N, K, ta, tb, "%.3f ms %.2f TF/s" % (ms.value, tf), flush=True)
out["gemm"] = res
ch = []
for n in (64, 128, 256, 512, 1000, 2000, 5000, 10000):
    Gm = rng.standard_normal((n, n))
    A = np.asfortranarray(Gm @ Gm.T / n + np.eye(n))
    info, ms = C.c_int32(), C.c_double()
    L.lrn_dbg_cholesky(n, dp(A), None, 0, C.byref(info), 5, C.byref(ms))
    tf = n ** 3 / 3 / (ms.value * 1e-3) / 1e12
    ch.append(dict(n=n, ms=ms.value, tflops=tf, info=info.value))
    print("chol", n, "%.3f ms %.2f TF/s info %d" % (ms.value, tf, info.value), flush=True)
out["cholesky"] = ch
sv = []
for m in (200, 801, 2000, 5000):
    A = np.asfortranarray(rng.standard_normal((m, m)))
    UD, V, sg = np.asfortranarray(np.zeros((m, m))), np.asfortranarray(np.zeros((m, m))), np.zeros(m)
    sw, ms = C.c_int32(), C.c_double()
    t = time.time()
    L.lrn_dbg_svd(m, dp(A), dp(UD), dp(V), dp(sg), 0.0, C.byref(sw), C.byref(ms))
    ref = np.linalg.svd(A, compute_uv=False)
    sv.append(dict(m=m, ms=ms.value, sweeps=sw.value, maxrel=float(np.max(np.abs(sg - ref) / ref))))
    print("svd", m, "%.1f ms sweeps %d maxrel %.2e" % (ms.value, sw.value, sv[-1]["maxrel"]), flush=True)
out["svd"] = sv
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
