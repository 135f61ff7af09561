"""Shared helpers of the double-double LP tests (tests/test_dd_cpu.py, tests/test_gpu_dd.py)."""
import numpy as np
import scipy.sparse as sp


def random_lp(n, nlin, seed, density=0.5):
    """Strictly feasible primal-dual pair in the reference's LP-block form (src/predictor_corrector.jl:19-20):
    C x = b, x >= 0 and C' y + s = d, s >= 0."""
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
    b = M @ xs
    d = M.T @ ys + ss
    return dict(n=n, msizes=[], A=[], b=b, b_const=0.0, C_lin=sp.csc_matrix(M), d_lin=d)
