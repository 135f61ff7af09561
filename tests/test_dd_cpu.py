"""CPU tests of the double-double LP path: the dd arithmetic the kernels use (host instantiation of
csrc/dd.cuh against __float128), and the pinning of the extended-precision oracle (oracle/dd_lp_oracle.py): at 53 bits it must
reproduce the Float64 oracle, which is itself pinned to the reference's examples; at 106 bits examples/k.jl:29-38."""
import os
import shutil
import subprocess

import mpmath as mp
import numpy as np
import pytest

import jump_examples as je
from dd_common import random_lp
from oracle import dd_lp_oracle as ddo
from oracle import loraine_oracle as lo
from oracle import sdpa_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dd_arithmetic_against_float128(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "dd_arith")
    cuda_inc = "/usr/local/cuda/include"
    r = subprocess.run([gxx, "-O2", "-ffp-contract=off", "-I" + cuda_inc, os.path.join(ROOT, "tests", "dd_arith_host.cpp"), "-o", exe,
                        "-lquadmath"], capture_output=True, text=True)
    if r.returncode != 0 and "quadmath" in r.stderr:
        pytest.skip("libquadmath not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout


def _models(spec):
    raw = sdpa_io.RawProblem(**je.fields(spec)) if "max_sense" in spec else sdpa_io.RawProblem(**spec)
    return raw, lo.prepare_model(raw, 0, 8)


@pytest.mark.parametrize("case", ["k", "random"])
def test_oracle_at_53_bits_is_the_float64_oracle(case):
    spec = je.ex_k_lp() if case == "k" else random_lp(6, 15, 1)
    raw, md = _models(spec)
    s64 = lo.solve_raw(raw, dict(kit=0, verb=0))
    s = ddo.solve(md.C_lin, md.d_lin, md.b, dict(eDIMACS=1e-7), prec=53)
    assert s64.status == 1 and s.status == 1
    assert s.iter == s64.iter
    assert abs(float(s.DIMACS_error) - s64.DIMACS_error) <= 1e-6 * s64.DIMACS_error
    assert np.allclose([float(v) for v in s.y], s64.y, rtol=1e-10, atol=1e-12)
    assert np.allclose([float(v) for v in s.X], s64.X_lin, rtol=1e-8, atol=1e-12)


def test_k_lp_float64x2():
    """examples/k.jl:8-38 with Optimizer{Float64x2}: objective 4, x = 2, shadow prices 0 and 2 -- here to 1e-24."""
    spec = je.ex_k_lp()
    raw, md = _models(spec)
    for promote_all in (False, True):
        s = ddo.solve(md.C_lin, md.d_lin, md.b, dict(eDIMACS=1e-25), prec=106, promote_all=promote_all)
        assert s.status == 1 and s.iter <= 30
        assert abs(s.y[0] * 2 - 4) <= mp.mpf(1e-24)            # objective = b'y with b = 2 (max sense)
        assert abs(s.y[0] - 2) <= mp.mpf(1e-24)
        assert abs(s.X[0]) <= mp.mpf(1e-24) and abs(s.X[1] - 2) <= mp.mpf(1e-24)
        assert s.DIMACS_error < mp.mpf(1e-25)


def test_first_iteration_is_float64_in_the_reference():
    """`ones(dd,1)` / `zeros(n,1)` are Float64 arrays (src/initial_point.jl:22,58,70): the faithful variant forms the first
    residuals in Float64, the all-T variant (what the CUDA path does) differs from it by Float64 rounding only."""
    spec = random_lp(5, 12, 3)
    raw, md = _models(spec)
    a = ddo.solve(md.C_lin, md.d_lin, md.b, dict(eDIMACS=1e-25), prec=106, promote_all=False, max_iters=1)
    b = ddo.solve(md.C_lin, md.d_lin, md.b, dict(eDIMACS=1e-25), prec=106, promote_all=True, max_iters=1)
    assert all(isinstance(v, float) for v in a.h_pred) and all(isinstance(v, mp.mpf) for v in b.h_pred)
    rel = max(abs(u - v) / (1 + abs(v)) for u, v in zip(a.y, b.y))
    assert 0 <= rel < 1e-13
    fa = ddo.solve(md.C_lin, md.d_lin, md.b, dict(eDIMACS=1e-25), prec=106, promote_all=False)
    fb = ddo.solve(md.C_lin, md.d_lin, md.b, dict(eDIMACS=1e-25), prec=106, promote_all=True)
    assert fa.status == fb.status == 1 and abs(fa.iter - fb.iter) <= 1
    assert max(abs(u - v) for u, v in zip(fa.y, fb.y)) < mp.mpf(1e-22)
