"""Seeded synthetic SDP generators for the benchmark configurations of BASELINE.json / SURVEY.md 8(d).

Every generator returns SDPA arrays `(n, blocksizes, c, body[k, blk, i, j, v])` (1-based blk/i/j, upper triangle),
i.e. exactly what a `.dat-s` file holds, so the same instance can feed the CUDA library, the oracle and -- written
with `model.write_sdpa` -- the Julia reference.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _torus_edges(rows, cols):
    idx = np.arange(rows * cols).reshape(rows, cols)
    right = np.stack([idx.ravel(), np.roll(idx, -1, axis=1).ravel()], axis=1)
    down = np.stack([idx.ravel(), np.roll(idx, -1, axis=0).ravel()], axis=1)
    e = np.concatenate([right, down])
    e.sort(axis=1)
    e = np.unique(e, axis=0)
    return e[e[:, 0] != e[:, 1]]


def maxcut_torus(rows=50, cols=100, seed=5000):
    """C2: max-cut relaxation in the layout of examples/data/maxG11.dat-s: min sum x_k  s.t. Diag(x) - L/4 >= 0,
    4-regular torus grid with +-1 weights.  F_k = e_k e_k' (rank one -> datarank = -1)."""
    N = rows * cols
    e = _torus_edges(rows, cols)
    rng = np.random.default_rng(seed)
    w = rng.choice([-1.0, 1.0], size=e.shape[0])
    deg = np.zeros(N)
    np.add.at(deg, e[:, 0], w)
    np.add.at(deg, e[:, 1], w)
    body = []
    nzd = np.nonzero(deg)[0]
    body.append(np.column_stack([np.zeros(nzd.size), np.ones(nzd.size), nzd + 1, nzd + 1, deg[nzd] / 4]))
    body.append(np.column_stack([np.zeros(e.shape[0]), np.ones(e.shape[0]), e[:, 0] + 1, e[:, 1] + 1, -w / 4]))
    k = np.arange(1, N + 1)
    body.append(np.column_stack([k, np.ones(N), k, k, np.ones(N)]))
    return N, [N], np.ones(N), np.concatenate(body)


def theta_torus(rows=20, cols=40):
    """C3: Lovasz theta in the layout of examples/data/thetaG11.dat-s (m = N+1, n_var = N + 1 + |E|).  Optimum N/2 for an
    even x even torus."""
    N = rows * cols
    e = _torus_edges(rows, cols)
    m = N + 1
    n = m + e.shape[0]
    i = np.arange(1, N + 1)
    body = [np.column_stack([np.zeros(N), np.ones(N), i, i, 0.5 * np.ones(N)]),
            np.column_stack([np.zeros(N), np.ones(N), i, m * np.ones(N), 0.25 * np.ones(N)])]
    k = np.arange(1, m + 1)
    body.append(np.column_stack([k, np.ones(m), k, k, np.ones(m)]))
    for t, (a, b) in enumerate(e):
        kk = m + 1 + t
        a1, b1 = a + 1, b + 1
        for (p, q) in ((a1, a1), (a1, b1), (a1, m), (b1, b1), (b1, m), (m, m)):
            body.append(np.array([[kk, 1, p, q, 1.0]]))
    return n, [m], np.ones(n), np.concatenate(body)


def multiblock_lp(nblocks=50, m=200, nlin=2000, seed=50):
    """C4: dual-form correlation-style problem: Diag(y_b) >= R_b for `nblocks` blocks of side m, plus an LP block of `nlin`
    rows with three +-1 coefficients each, strictly feasible at y0 = lambda_max(R_b) + 1."""
    n = nblocks * m
    body = []
    y0 = np.zeros(n)
    for b in range(nblocks):
        rng = np.random.default_rng(seed + b)
        Nn = rng.standard_normal((m, m))
        R = 0.5 * (Nn + Nn.T) / np.sqrt(m)
        y0[b * m:(b + 1) * m] = np.linalg.eigvalsh(R)[-1] + 1.0
        iu, ju = np.triu_indices(m)
        body.append(np.column_stack([np.zeros(iu.size), (b + 1) * np.ones(iu.size), iu + 1, ju + 1, R[iu, ju]]))
        p = np.arange(1, m + 1)
        body.append(np.column_stack([b * m + p, (b + 1) * np.ones(m), p, p, np.ones(m)]))
    bs = [m] * nblocks
    if nlin > 0:
        rng = np.random.default_rng(seed - 1)
        lp = nblocks + 1
        for r in range(nlin):
            vars_ = rng.choice(n, size=3, replace=False)
            sg = rng.choice([-1.0, 1.0], size=3)
            for v_, s_ in zip(vars_, sg):
                body.append(np.array([[v_ + 1, lp, r + 1, r + 1, s_]]))
            body.append(np.array([[0, lp, r + 1, r + 1, float(sg @ y0[vars_]) - 1.0]]))
        bs = bs + [-nlin]
    return n, bs, np.ones(n), np.concatenate(body)


def large_schur(m=1000, n=40000, seed=40000):
    """C5: one block of side m with n sparse symmetric constraint matrices: 90 % have one off-diagonal pair (2 stored
    entries), 10 % are clique patterns on 3-5 indices (9-25 stored entries); F_0 = -(I + small noise);
    c_k = <F_k, X*> for a random SPD X* so that the problem is strictly feasible on both sides."""
    rng = np.random.default_rng(seed)
    Nn = rng.standard_normal((m, m))
    Xs = Nn @ Nn.T / m + np.eye(m)
    body, c = [], np.zeros(n)
    E = 0.01 * rng.standard_normal((m, m))
    F0 = -(np.eye(m) + 0.5 * (E + E.T) / np.sqrt(m))
    iu, ju = np.triu_indices(m)
    body.append(np.column_stack([np.zeros(iu.size), np.ones(iu.size), iu + 1, ju + 1, F0[iu, ju]]))
    ks, ii, jj, vv = [], [], [], []
    n_clique = n // 10
    kinds = np.zeros(n, dtype=bool)
    kinds[rng.choice(n, size=n_clique, replace=False)] = True
    # distinct off-diagonal positions (otherwise the constraint matrices would be linearly dependent)
    npairs = m * (m - 1) // 2
    pair_ids = rng.choice(npairs, size=n - n_clique, replace=False)
    iu1, ju1 = np.triu_indices(m, k=1)
    next_pair = 0
    for k in range(n):
        if not kinds[k]:
            a, b = int(iu1[pair_ids[next_pair]]), int(ju1[pair_ids[next_pair]])
            next_pair += 1
            v = rng.standard_normal()
            ks.append(k + 1); ii.append(a + 1); jj.append(b + 1); vv.append(v)
            c[k] = 2 * v * Xs[a, b]
        else:
            s = rng.integers(3, 6)
            idx = np.sort(rng.choice(m, size=s, replace=False))
            V = rng.standard_normal((s, s))
            V = 0.5 * (V + V.T)
            for x in range(s):
                for y in range(x, s):
                    ks.append(k + 1); ii.append(idx[x] + 1); jj.append(idx[y] + 1); vv.append(V[x, y])
            c[k] = float(np.sum(V * Xs[np.ix_(idx, idx)]))
    body.append(np.column_stack([ks, np.ones(len(ks)), ii, jj, vv]))
    return n, [m], c, np.concatenate(body)



def random_lp(n, nlin, seed, density=0.5):
    """Strictly feasible primal-dual LP pair in the reference's LP-block form (src/predictor_corrector.jl:19-20):
    C x = b, x >= 0 and C' y + s = d, s >= 0 (no PSD block).  Returns the RawProblem fields."""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, nlin)) * (rng.random((n, nlin)) < density)
    for i in range(n):                                   # no empty row / column
        M[i, rng.integers(nlin)] += 1.0 + rng.random()
    for k in range(nlin):
        if not M[:, k].any():
            M[rng.integers(n), k] = 1.0
    xs = 0.5 + rng.random(nlin)
    ys = rng.standard_normal(n)
    ss = 0.5 + rng.random(nlin)
    return dict(n=n, msizes=[], A=[], b=M @ xs, b_const=0.0, C_lin=sp.csc_matrix(M), d_lin=M.T @ ys + ss)


CONFIGS = {
    "C2": dict(gen=lambda: maxcut_torus(50, 100, 5000),
               options=dict(kit=0, datarank=-1, initpoint=1, eDIMACS=1e-6)),
    "C2-mini": dict(gen=lambda: maxcut_torus(8, 12, 96), options=dict(kit=0, datarank=-1, initpoint=1, eDIMACS=1e-6)),
    "C3": dict(gen=lambda: theta_torus(20, 40),
               options=dict(kit=1, preconditioner=1, erank=1, aamat=2, tol_cg=1e-2, tol_cg_up=0.5, tol_cg_min=1e-6,
                            eDIMACS=1e-5, initpoint=1)),
    "C3-mini": dict(gen=lambda: theta_torus(6, 8),
                    options=dict(kit=1, preconditioner=1, erank=1, aamat=2, tol_cg=1e-2, tol_cg_up=0.5, tol_cg_min=1e-6,
                                 eDIMACS=1e-5, initpoint=1)),
    "C4": dict(gen=lambda: multiblock_lp(50, 200, 2000, 50), options=dict(kit=0, datarank=0, initpoint=1, eDIMACS=1e-6)),
    "C4-mini": dict(gen=lambda: multiblock_lp(5, 20, 40, 50), options=dict(kit=0, datarank=0, initpoint=1, eDIMACS=1e-6)),
    "C5": dict(gen=lambda: large_schur(1000, 40000, 40000),
               options=dict(kit=0, datarank=0, datasparsity=8, initpoint=1, eDIMACS=1e-6)),
    "C5-mini": dict(gen=lambda: large_schur(30, 200, 40000),
                    options=dict(kit=0, datarank=0, datasparsity=8, initpoint=1, eDIMACS=1e-6)),
}
