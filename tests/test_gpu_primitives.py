"""GPU parity tests of the dense primitives (through the debug C ABI) against NumPy/LAPACK on the same seeded inputs.
Tolerances are stated per test (Float64 path)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(pkg):
    from loraine_jl_b200 import _lib
    lib = _lib.lib()
    i32, dbl = C.c_int32, C.c_double
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.lrn_dbg_gemm.argtypes = [i32, i32, i32, i32, i32, dbl, pd, pd, dbl, pd, i32, i32, pd, i32, i32, pd]
    lib.lrn_dbg_cholesky.argtypes = [i32, pd, pd, i32, pi, i32, pd]
    lib.lrn_dbg_eig_small.argtypes = [i32, pd, pd, pd, i32]
    lib.lrn_dbg_svd.argtypes = [i32, pd, pd, pd, pd, dbl, pi, pd]
    lib.lrn_dbg_lanczos.argtypes = [i32, pd, i32, dbl, pd, pd, pd, pd, pi, pi]
    lib.lrn_dbg_peak.argtypes = [i32, pd]
    lib.lrn_dbg_batched_lambda_min.argtypes = [i32, i32, pd, pd]
    return lib


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def F(a):
    return np.asfortranarray(a, dtype=np.float64)


def run_gemm(L, A, B, Cm, ta, tb, alpha, beta, mode=0, lower=0, cs=None, misalign=0):
    M, N = Cm.shape
    K = A.shape[0] if ta else A.shape[1]
    A, B, out = F(A), F(B), F(Cm.copy())
    rc = L.lrn_dbg_gemm(M, N, K, ta, tb, alpha, dp(A), dp(B), beta, dp(out), mode, lower, dp(cs) if cs is not None else None,
                        misalign, 0, None)
    assert rc == 0
    return out


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 64, 64), (50, 50, 50), (130, 70, 33), (200, 200, 200),
                                   (257, 511, 129), (801, 801, 801), (1024, 1536, 512), (2000, 64, 64)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_matches_numpy(L, M, N, K, ta, tb):
    rng = np.random.default_rng(M * 1000 + N * 10 + K + ta * 2 + tb)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    want = 1.5 * (A.T if ta else A) @ (B.T if tb else B) - 0.5 * C0
    got = run_gemm(L, A, B, C0, ta, tb, 1.5, -0.5)
    # |err| <= K * eps * |A||B| ; relative Frobenius 1e-13 is generous for K <= 801
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want) + 1e-300


@pytest.mark.parametrize("M,N,K,lower,mode", [(1024, 1024, 64, 0, 0), (1100, 1300, 96, 0, 0), (2000, 1030, 512, 0, 0), (1500, 1500, 256, 1, 0),
                                              (1280, 1280, 160, 1, 1), (1300, 1300, 32 * 7, 1, 1)])
def test_gemm_tma_bulk_path(L, M, N, K, lower, mode):
    """A B^T products with K % 32 == 0 and M, N >= 1024 run on the TMA-fed kernel (cp.async.bulk + mbarrier producer warp) for
    the full 128 x 128 tiles and on the cp.async kernel for the edge strips; both regions are checked against NumPy."""
    rng = np.random.default_rng(M + N + K)
    A, B = rng.standard_normal((M, K)), rng.standard_normal((N, K))
    if lower:
        B = A[:N]
    C0 = rng.standard_normal((M, N))
    P = A @ B.T
    want = C0 + (P ** 2 if mode else -0.75 * P)
    got = run_gemm(L, A, B, C0, 0, 1, 1.0 if mode else -0.75, 1.0, mode=mode, lower=lower)
    if lower:
        il = np.tril_indices(min(M, N))
        assert np.linalg.norm(got[il] - want[il]) <= 1e-13 * np.linalg.norm(want[il])
    else:
        assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)


@pytest.mark.parametrize("M,N,K,tb,lower,beta", [(1024, 1024, 64, 0, 0, 0.0), (1100, 1300, 104, 0, 0, 1.0), (1300, 1100, 1000, 0, 0, 0.5),
                                                 (1500, 1500, 1500, 0, 1, 0.0), (1500, 1500, 1500, 1, 1, 0.0), (1030, 2000, 77, 1, 0, 1.0)])
def test_gemm_tma_bulk_nn_and_k_remainder(L, M, N, K, tb, lower, beta):
    """A B products (K-major B operand moved as 256 B segments) and K % 32 != 0 (remainder added by the generic kernel)."""
    rng = np.random.default_rng(M + 3 * N + 7 * K + tb)
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K)) if tb else rng.standard_normal((K, N))
    if lower:                                     # symmetric product: B = A' (or A for the transposed form)
        B = A[:N] if tb else A[:N].T.copy()
    C0 = rng.standard_normal((M, N))
    P = A @ (B.T if tb else B)
    want = beta * C0 + 1.25 * P
    got = run_gemm(L, A, B, C0, 0, tb, 1.25, beta, mode=0, lower=lower)
    if lower:
        il = np.tril_indices(min(M, N))
        assert np.linalg.norm(got[il] - want[il]) <= 1e-13 * np.linalg.norm(want[il])
    else:
        assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)


def test_gemm_unaligned_colscale_beta0_nan_safe(L):
    rng = np.random.default_rng(5)
    A, B = rng.standard_normal((123, 77)), rng.standard_normal((77, 95))
    cs = rng.standard_normal(95)
    C0 = np.full((123, 95), np.nan)             # beta = 0 must not read C
    got = run_gemm(L, A, B, C0, 0, 0, 1.0, 0.0, cs=cs, misalign=1)
    want = (A @ B) * cs[None, :]
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)


def test_gemm_square_epilogue_lower(L):
    """rank-one Schur epilogue: C += (A A').^2 on the lower triangle (src/makeBBBB.jl:10-14)."""
    rng = np.random.default_rng(6)
    A = rng.standard_normal((300, 90))
    C0 = rng.standard_normal((300, 300))
    got = run_gemm(L, A, A, C0, 0, 1, 1.0, 1.0, mode=1, lower=1)
    want = C0 + (A @ A.T) ** 2
    il = np.tril_indices(300)
    assert np.linalg.norm(got[il] - want[il]) <= 1e-13 * np.linalg.norm(want[il])


@pytest.mark.parametrize("n", [1, 5, 64, 65, 100, 200, 513, 1000, 2500])
def test_cholesky_and_solves(L, n):
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n))
    A = G @ G.T / n + np.eye(n)
    b = rng.standard_normal(n)
    Aio, x = F(A.copy()), b.copy()
    info = C.c_int32(-1)
    assert L.lrn_dbg_cholesky(n, dp(Aio), dp(x), 3, C.byref(info), 0, None) == 0
    assert info.value == 0
    Lref = np.linalg.cholesky(A)
    assert np.linalg.norm(Aio - Lref) <= 1e-12 * np.linalg.norm(Lref)
    xref = np.linalg.solve(A, b)
    assert np.linalg.norm(x - xref) <= 1e-11 * np.linalg.norm(xref)
    for which in (1, 2):
        x = b.copy()
        Aio = F(A.copy())
        assert L.lrn_dbg_cholesky(n, dp(Aio), dp(x), which, C.byref(info), 0, None) == 0
        ref = np.linalg.solve(Lref if which == 1 else Lref.T, b)
        assert np.linalg.norm(x - ref) <= 1e-11 * np.linalg.norm(ref)


def test_cholesky_reports_first_bad_pivot(L):
    n = 150
    A = np.eye(n)
    A[100, 100] = -1.0
    Aio = F(A)
    info = C.c_int32(0)
    assert L.lrn_dbg_cholesky(n, dp(Aio), None, 0, C.byref(info), 0, None) == 0
    assert info.value == 101                        # LAPACK: leading minor of order 101 is not positive definite


@pytest.mark.parametrize("n", [1, 2, 3, 10, 33, 50, 64])
def test_jacobi_eig_small(L, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A = (A + A.T) / 2
    ev, V = np.zeros(n), F(np.zeros((n, n)))
    assert L.lrn_dbg_eig_small(n, dp(F(A)), dp(ev), dp(V), 0) == 0
    ref = np.linalg.eigvalsh(A)[::-1]
    assert np.max(np.abs(ev - ref)) <= 1e-13 * max(1.0, np.abs(ref).max())
    assert np.linalg.norm(V.T @ V - np.eye(n)) <= 1e-13 * n
    assert np.linalg.norm(A @ V - V * ev[None, :]) <= 1e-12 * max(1.0, np.linalg.norm(A))


@pytest.mark.parametrize("m,cond", [(3, 1e8), (50, 1e8), (64, 1e6), (65, 1e3), (130, 1e3), (200, 1e2), (801, 30.0), (1500, 5.0)])
def test_block_jacobi_svd(L, m, cond):
    """L_S' L_X of interior-point iterates is well conditioned (cond ~ 1..1e2, measured); single-pair sizes are also
    exercised with strongly graded spectra."""
    rng = np.random.default_rng(m)
    U0, _ = np.linalg.qr(rng.standard_normal((m, m)))
    V0, _ = np.linalg.qr(rng.standard_normal((m, m)))
    sv = np.logspace(0, -np.log10(cond), m) * 37.0
    A = (U0 * sv[None, :]) @ V0.T
    UD, V, sg = F(np.zeros((m, m))), F(np.zeros((m, m))), np.zeros(m)
    sweeps, ms = C.c_int32(0), C.c_double(0)
    assert L.lrn_dbg_svd(m, dp(F(A)), dp(UD), dp(V), dp(sg), 0.0, C.byref(sweeps), C.byref(ms)) == 0
    ref = np.linalg.svd(A, compute_uv=False)
    assert np.max(np.abs(sg - ref) / ref) <= 1e-8             # relative accuracy of every singular value
    assert np.linalg.norm(V.T @ V - np.eye(m)) <= 1e-12 * m
    assert np.linalg.norm(A @ V - UD) <= 1e-12 * np.linalg.norm(A)
    Un = UD / sg[None, :]
    assert np.linalg.norm(Un.T @ Un - np.eye(m)) <= 1e-8 * m   # left vectors of tiny singular values are less accurate
    assert 1 <= sweeps.value <= 25


@pytest.mark.parametrize("m", [5, 64, 257, 700, 1500])
def test_gemm_triangular_operands(L, m):
    """CC = L_S' L_X (src/prepare_W.jl:39): both factors lower triangular, tiles start their K loop at max(m0, n0)."""
    rng = np.random.default_rng(m)
    LS = np.tril(rng.standard_normal((m, m)))
    LX = np.tril(rng.standard_normal((m, m)))
    Cm = F(np.zeros((m, m)))
    assert L.lrn_dbg_gemm(m, m, m, 1, 0, 1.0, dp(F(LS)), dp(F(LX)), 0.0, dp(Cm), 2, 0, None, 0, 0, None) == 0
    ref = LS.T @ LX
    assert np.max(np.abs(Cm - ref)) <= 1e-13 * m * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("m", [300, 1030, 2100])
def test_block_jacobi_svd_without_v(L, m):
    """The solver never accumulates V (G comes from triangular solves); m >= 1024 runs the TMA panel-rotation kernel and
    the recycled diagonal Gram blocks."""
    rng = np.random.default_rng(m)
    U0, _ = np.linalg.qr(rng.standard_normal((m, m)))
    V0, _ = np.linalg.qr(rng.standard_normal((m, m)))
    sv = np.linspace(1.0, 5.0, m)[::-1]
    A = (U0 * sv[None, :]) @ V0.T
    UD, sg = F(np.zeros((m, m))), np.zeros(m)
    sweeps, ms = C.c_int32(0), C.c_double(0)
    assert L.lrn_dbg_svd(m, dp(F(A)), dp(UD), None, dp(sg), 1e-10, C.byref(sweeps), C.byref(ms)) == 0
    assert np.max(np.abs(sg - sv) / sv) <= 1e-12
    Un = UD / sg[None, :]
    assert np.linalg.norm(Un.T @ Un - np.eye(m)) <= 1e-9 * m
    # A = U D V' with orthogonal V  <=>  A A' = (UD)(UD)'
    assert np.linalg.norm(A @ A.T - UD @ UD.T) <= 1e-12 * np.linalg.norm(A) ** 2
    assert 1 <= sweeps.value <= 25


@pytest.mark.parametrize("m", [10, 64, 100, 500, 1200])
def test_lanczos_extremes(L, m):
    rng = np.random.default_rng(m)
    Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    lam = np.concatenate([[-3.0], np.linspace(-1, 1, m - 3), [5.0, 9.0]]) if m > 3 else np.array([-1.0, 0.5, 2.0])[:m]
    T = (Q * lam[None, :]) @ Q.T
    T = (T + T.T) / 2
    lmin, lmax = C.c_double(), C.c_double()
    tv, tvec = np.zeros(2), F(np.zeros((m, 2)))
    it, conv = C.c_int32(), C.c_int32()
    assert L.lrn_dbg_lanczos(m, dp(F(T)), 2, 1e-11, C.byref(lmin), C.byref(lmax), dp(tv), dp(tvec), C.byref(it), C.byref(conv)) == 0
    assert conv.value == 1
    assert abs(lmin.value - lam.min()) <= 1e-9 * np.abs(lam).max()
    ref = np.sort(lam)[-2:]
    assert np.max(np.abs(tv - ref)) <= 1e-9 * np.abs(lam).max()
    assert np.linalg.norm(T @ tvec - tvec * tv[None, :]) <= 1e-7 * np.abs(lam).max()


@pytest.mark.parametrize("m,count", [(1, 3), (2, 2), (3, 4), (65, 5), (200, 7), (384, 3)])
def test_batched_lambda_min(L, m, count):
    """one CTA per matrix: Householder tridiagonalisation + Sturm multisection vs LAPACK eigvalsh"""
    rng = np.random.default_rng(m + count)
    mats = np.zeros((count, m, m))
    for z in range(count):
        A = rng.standard_normal((m, m))
        mats[z] = (A + A.T) / 2 - (z % 2) * 3.0 * np.eye(m)
    if count > 2 and m > 2:
        mats[2] = np.diag(np.linspace(-2.0, 5.0, m))            # already diagonal (all reflections skipped)
    flat = np.ascontiguousarray(np.stack([np.asfortranarray(a).ravel(order="F") for a in mats]))
    out = np.zeros(count)
    assert L.lrn_dbg_batched_lambda_min(count, m, dp(flat), dp(out)) == 0
    ref = np.array([np.linalg.eigvalsh(a)[0] for a in mats])
    scale = np.array([max(1.0, np.abs(np.linalg.eigvalsh(a)).max()) for a in mats])
    assert np.max(np.abs(out - ref) / scale) <= 1e-13
