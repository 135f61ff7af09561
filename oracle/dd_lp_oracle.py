"""TEST INFRASTRUCTURE (oracle, not product code): extended-precision restatement of Loraine.jl's interior-point iteration for
models WITHOUT semidefinite blocks, element type `Float64x2` (`Optimizer{Float64x2}`, reference README.md:37-54,
examples/k.jl:8).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

The reference runs this case through its generic Julia code with MultiFloats.jl numbers.  This file follows the same
statements with mpmath numbers of `prec` bits (Float64x2 carries ~106):

    initial_point / find_initial!      src/initial_point.jl:1-81        (plain Float64 arithmetic: `ones(dd,1)`, `zeros(n,1)`)
    find_mu                            src/Solvers.jl:480-494           (result stored in a `::T` field)
    prepare_W (LP part)                src/prepare_W.jl:86
    predictor                          src/predictor_corrector.jl:5-146 (`BBBB = zeros(T, n, n)`; `BBBB .+= C_lin spdiagm(...) C_lin'`)
    sigma_update                       src/predictor_corrector.jl:148-179
    corrector                          src/predictor_corrector.jl:181-246
    find_step_lin                      src/predictor_corrector.jl:329-364
    check_convergence                  src/Solvers.jl:496-568
    solve loop, myIPstep               src/Solvers.jl:304-361, :448-478

Element types follow Julia's promotion exactly as Python's does: a Python float stands for a Float64, an mpf for a T; float
op float stays Float64, float op mpf gives T.  So, as in the reference, the first iteration's residuals and right-hand side are
formed in Float64 (the iterate is still `ones`/`zeros` of Float64) and everything is T from the first Cholesky solve on.
`promote_all=True` converts the initial point to T instead (what the CUDA path does: it keeps every array in double-double
from the start); the two variants differ by Float64 rounding of the first iteration only.

Pinning: with prec = 53 mpmath's arithmetic is IEEE double arithmetic, and this file must reproduce the Float64 oracle
(oracle/loraine_oracle.py, itself pinned to the reference's examples) on the same LP -- tests/test_oracle_golden.py checks that,
and examples/k.jl:29-38 (objective 4, x = 2, shadow prices 0 and 2).
"""
from __future__ import annotations

import math

import mpmath as mp
import numpy as np
import scipy.sparse as sp


class DDLPState:
    pass


def _rows_cols(C_lin):
    Cc = sp.csc_matrix(C_lin)
    Cc.sum_duplicates()
    Cc.sort_indices()
    n, m = Cc.shape
    cols = [[(int(Cc.indices[e]), float(Cc.data[e])) for e in range(Cc.indptr[k], Cc.indptr[k + 1])] for k in range(m)]
    rows = [[] for _ in range(n)]
    for k in range(m):
        for i, v in cols[k]:
            rows[i].append((k, v))
    return rows, cols


def _C_times(rows, v):                 # C_lin * v
    out = []
    for row in rows:
        acc = 0.0
        for k, c in row:
            acc = acc + c * v[k]
        out.append(acc)
    return out


def _Ct_times(cols, v):                # C_lin' * v
    out = []
    for col in cols:
        acc = 0.0
        for i, c in col:
            acc = acc + c * v[i]
        out.append(acc)
    return out


def _dot(a, b):
    acc = 0.0
    for u, v in zip(a, b):
        acc = acc + u * v
    return acc


def _norm(a):
    return mp.sqrt(_dot(a, a))


def _T(v):
    return mp.mpf(v)


def cholesky_lower(H):
    """cholesky(Hermitian(BBBB, :L)) in T; raises ArithmeticError(k) at the first non-positive pivot (1-based k)."""
    n = len(H)
    L = [[mp.mpf(0)] * n for _ in range(n)]
    for j in range(n):
        d = H[j][j]
        for p in range(j):
            d = d - L[j][p] * L[j][p]
        if not d > 0:
            raise ArithmeticError(j + 1)
        L[j][j] = mp.sqrt(d)
        for i in range(j + 1, n):
            v = H[i][j]
            for p in range(j):
                v = v - L[i][p] * L[j][p]
            L[i][j] = v / L[j][j]
    return L


def _fwd(L, h):
    n = len(h)
    x = list(h)
    for i in range(n):
        v = x[i]
        for p in range(i):
            v = v - L[i][p] * x[p]
        x[i] = v / L[i][i]
    return x


def _bwd(L, h):
    n = len(h)
    x = list(h)
    for i in range(n - 1, -1, -1):
        v = x[i]
        for p in range(i + 1, n):
            v = v - L[p][i] * x[p]
        x[i] = v / L[i][i]
    return x


def _solve_direct(s, h):
    """`cholBBBB' \\ (cholBBBB \\ h)` (src/predictor_corrector.jl:89-90, :199); after a regularised retry cholBBBB is a `Cholesky`
    object whose adjoint is itself, so the expression evaluates H^-1 H^-1 h (reference quirk, kept as in loraine_oracle.py)."""
    x = _bwd(s.L, _fwd(s.L, h))
    if s.chol_is_factor_object:
        x = _bwd(s.L, _fwd(s.L, x))
    return x


def _steplen(mimi, tau):
    return 0.99 if mimi > -1e-6 else min(1, -tau / mimi)


def setup(C_lin, d_lin, b, options=None, prec=106, promote_all=False):
    o = dict(eDIMACS=1e-7, maxit=100, initpoint=0, verb=0)
    o.update(options or {})
    mp.mp.prec = prec
    s = DDLPState()
    s.prec, s.promote_all = prec, promote_all
    s.rows, s.cols = _rows_cols(C_lin)
    s.n, s.nlin = len(s.rows), len(s.cols)
    s.b = [float(v) for v in np.asarray(b, float).ravel()]
    s.d = [float(v) for v in np.asarray(d_lin, float).ravel()]
    s.eDIMACS = _T(o["eDIMACS"])
    s.maxit, s.initpoint, s.verb = int(o["maxit"]), int(o["initpoint"]), int(o["verb"])
    s.hooks = {}
    s.regcount = 0
    s.history = []
    return s


def initial_point(s):
    """src/initial_point.jl:1-81 -- Float64 arithmetic (`zeros(n,1)`, `Epss * ones(dd,1)`)."""
    n, dd = s.n, s.nlin
    s.y = [0.0] * n
    b2 = [1 + abs(v) for v in s.b]
    rown = [math.sqrt(sum(c * c for _, c in row)) for row in s.rows]
    if s.initpoint == 0:
        Epss, Etaa = 1.0, 1.0
    else:
        Epss = max(1.0, max(b2[j] / (1 + rown[j]) for j in range(n)))
        mf = max(max(rown), math.sqrt(sum(v * v for v in s.d))) / math.sqrt(dd)
        Etaa = max(1.0, mf)
    s.X = [Epss] * dd
    s.S = [Etaa] * dd
    if s.promote_all:
        s.X = [_T(v) for v in s.X]
        s.S = [_T(v) for v in s.S]
        s.y = [_T(v) for v in s.y]
    s.S_inv = [1 / v for v in s.S]
    s.sigma = _T(3)
    s.tau = _T(0.95)
    s.expon = _T(3.0)
    s.DIMACS_error = _T(1.0)
    s.iter = 0
    s.status = 0
    s.alpha = s.beta = _T(1)
    s.RNT = [0.0] * dd


def find_mu(s):
    s.mu = _T(_dot(s.X, s.S)) / (0 + s.nlin)
    return s.mu


def prepare_W(s):
    s.Si = [1.0 / v for v in s.S]


def residuals(s):
    cx = _C_times(s.rows, s.X)
    s.Rp = [bi - v for bi, v in zip(s.b, cx)]
    cty = _Ct_times(s.cols, s.y)
    s.Rd = [di - si - v for di, si, v in zip(s.d, s.S, cty)]


def schur(s):
    """BBBB = zeros(T, n, n); BBBB .+= C_lin * spdiagm(X_lin .* S_lin_inv) * C_lin'; Hermitian(BBBB, :L)."""
    n = s.n
    w = [x * si for x, si in zip(s.X, s.S_inv)]
    H = [[0.0] * n for _ in range(n)]
    for k, col in enumerate(s.cols):
        for i, ci in col:
            t = ci * w[k]
            for j, cj in col:
                if j <= i:
                    H[i][j] = H[i][j] + t * cj
    Ht = [[mp.mpf(0)] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1):
            Ht[i][j] = Ht[j][i] = _T(H[i][j])
    return Ht


def rhs_predictor(s):
    t = [(x * si) * rd + x for x, si, rd in zip(s.X, s.Si, s.Rd)]
    ct = _C_times(s.rows, t)
    return [r + v for r, v in zip(s.Rp, ct)]


def rhs_corrector(s):
    sm = s.sigma * s.mu
    t = []
    for x, si, rd, dx, ds in zip(s.X, s.Si, s.Rd, s.delX, s.delS):
        tmp = (dx * ds) * si - sm * si
        t.append((x * si) * rd + x + tmp)
    ct = _C_times(s.rows, t)
    return [r + v for r, v in zip(s.Rp, ct)]


def find_step_lin(s):
    cty = _Ct_times(s.cols, s.dely)
    s.delS = [rd - v for rd, v in zip(s.Rd, cty)]
    if s.predict:
        s.delX = [-x - x * si * ds for x, si, ds in zip(s.X, s.Si, s.delS)]
    else:
        sm = s.sigma * s.mu
        s.delX = [-x - x * si * ds + sm * si + rnt for x, si, ds, rnt in zip(s.X, s.Si, s.delS, s.RNT)]
    s.mimiX = min(dx / x for dx, x in zip(s.delX, s.X))
    s.mimiS = min(ds / sv for ds, sv in zip(s.delS, s.S))
    s.alpha = _T(_steplen(s.mimiX, s.tau))
    s.beta = _T(_steplen(s.mimiS, s.tau))
    if s.predict:
        s.Xn = [x + s.alpha * dx for x, dx in zip(s.X, s.delX)]
        s.Sn = [sv + s.beta * ds for sv, ds in zip(s.S, s.delS)]
        s.RNT = [-(dx * ds) * si for dx, ds, si in zip(s.delX, s.delS, s.Si)]
    else:
        s.y = [yv + s.beta * dv for yv, dv in zip(s.y, s.dely)]
        s.X = [x + s.alpha * dx for x, dx in zip(s.X, s.delX)]
        s.S = [sv + s.beta * ds for sv, ds in zip(s.S, s.delS)]
        s.S_inv = [1 / v for v in s.S]


def predictor(s):
    s.predict = True
    residuals(s)
    H = schur(s)
    s.H = H
    h = rhs_predictor(s)
    s.h_pred = h
    try:
        s.L = cholesky_lower(H)
        s.chol_is_factor_object = False
    except ArithmeticError:
        icount = 0
        s.regcount += 1
        if s.regcount > 5:
            s.status = 3
            return
        while True:
            try:
                s.L = cholesky_lower(H)
                break
            except ArithmeticError:
                for i in range(s.n):
                    H[i][i] = H[i][i] + 1e-4
                icount += 1
                if icount > 1000:
                    s.status = 3
                    return
        s.chol_is_factor_object = True
    s.dely = _solve_direct(s, h)
    s.dely_pred = list(s.dely)
    find_step_lin(s)


def sigma_update(s):
    """src/predictor_corrector.jl:148-179; the ratio and the power are evaluated in Float64 (`convert(Float64, tmp12)`,
    `Float64(solver.mu)`, `Float64(expon_used)`, :173-175) and the result is stored in the `::T` field."""
    step_pred = min(s.alpha, s.beta)
    if s.mu > 1e-6:
        if step_pred < 1 / math.sqrt(3):
            expon_used = 1.0
        else:
            expon_used = max(s.expon, _T(3) * step_pred ** 2)
    else:
        expon_used = max(1, min(s.expon, _T(3) * step_pred ** 2))
    tmp12 = float(_dot(s.Xn, s.Sn) / (0 + s.nlin))         # btrace over zero PSD blocks is 0 (not negative)
    s.sigma = _T(min(1.0, (tmp12 / float(s.mu)) ** float(expon_used)))


def corrector(s):
    s.predict = False
    h = rhs_corrector(s)
    s.h_corr = h
    s.dely = _solve_direct(s, h)
    find_step_lin(s)


def check_convergence(s):
    nb = math.sqrt(sum(v * v for v in s.b))
    nd = math.sqrt(sum(v * v for v in s.d))
    by = _dot(s.b, s.y)
    dx = _dot(s.d, s.X)
    s.err1 = _norm(s.Rp) / (1 + nb)
    s.err2 = max(0, -min(s.X) / (1 + nb))
    s.err3 = _norm(s.Rd) / (1 + nd)
    s.err4 = max(0, -min(s.S) / (1 + nd))
    s.err5 = (dx - by) / (1 + abs(by))
    s.err6 = _dot(s.S, s.X) / (1 + abs(dx) + abs(by))
    s.DIMACS_error = s.err2 + s.err3 + s.err4 + abs(s.err5) + s.err6          # nlmi = 0: err1 is left out (src/Solvers.jl:521)
    s.by, s.dx = by, dx
    if s.DIMACS_error < s.eDIMACS:
        s.status = 1
    if s.DIMACS_error > 1e55:
        s.status = 2
    elif abs(by) > 1e55:
        s.status = 3


def myIPstep(s):
    s.iter += 1
    if s.iter > s.maxit:
        s.status = 4
    find_mu(s)
    prepare_W(s)
    predictor(s)
    if s.status == 3:
        return
    sigma_update(s)
    corrector(s)


def solve(C_lin, d_lin, b, options=None, prec=106, promote_all=False, max_iters=None, hooks=None):
    s = setup(C_lin, d_lin, b, options, prec, promote_all)
    s.hooks = hooks or {}
    initial_point(s)
    while s.status == 0:
        myIPstep(s)
        check_convergence(s)
        s.history.append(dict(iter=s.iter, mu=s.mu, sigma=s.sigma, alpha=s.alpha, beta=s.beta, DIMACS_error=s.DIMACS_error, by=s.by))
        if "iteration" in s.hooks:
            s.hooks["iteration"](s)
        if max_iters is not None and s.iter >= max_iters:
            break
    return s
