"""Host-side model preparation (the library's INPUTS).  Mirrors src/model.jl:34-229 and the triplet builder of
src/MOI_wrapper.jl:142-223, plus an SDPA (.dat-s) loader equivalent to examples/solve_sdpa.jl:14-34.

Everything here is one-off preprocessing on the host; the per-iteration hot path lives in
the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


@dataclass
class RawProblem:
    """What `MOI.copy_to` (src/MOI_wrapper.jl:142-232) holds before `_prepare_A`: per PSD block the triplets of
    A[i, k], k = 0..n (k = 0 is A[i,1] = -constants), raw signs, BOTH triangles stored; b, b_const, C_lin, d_lin."""
    n: int
    msizes: list
    A: list                      # per block dict(k, p, q, v) with 0-based p, q
    b: np.ndarray
    b_const: float = 0.0
    C_lin: sp.csc_matrix | None = None     # n x nlin
    d_lin: np.ndarray = field(default_factory=lambda: np.zeros(0))

    @property
    def nlmi(self):
        return len(self.msizes)

    @property
    def nlin(self):
        return int(np.asarray(self.d_lin).shape[0])


def _tokens(path):
    with open(path) as f:
        for ln in f:
            s = ln.strip()
            if not s or s[0] in '"*':
                continue
            yield s


def read_sdpa(path):
    """SDPA sparse format -> (n, blocksizes, c, entries[k, blk, i, j, v]).  Negative block size = diagonal (LP) block."""
    clean = str.maketrans("{}(),", "     ")
    it = _tokens(path)
    n = int(float(next(it).translate(clean).split()[0]))
    nblocks = int(float(next(it).translate(clean).split()[0]))
    bs = []
    while len(bs) < nblocks:
        bs += [int(float(t)) for t in next(it).translate(clean).split()]
    bs = bs[:nblocks]
    c = []
    while len(c) < n:
        c += [float(t.replace("D", "e").replace("d", "e")) for t in next(it).translate(clean).split()]
    rows = []
    for s in it:
        t = s.translate(clean).split()
        if len(t) >= 5:
            rows.append([float(x.replace("D", "e").replace("d", "e")) for x in t[:5]])
    body = np.array(rows, dtype=np.float64).reshape(-1, 5)
    return n, bs, np.array(c[:n]), body


def write_sdpa(path, n, bs, c, body):
    with open(path, "w") as f:
        f.write(f"{n}\n{len(bs)}\n{' '.join(str(int(b)) for b in bs)}\n")
        f.write(" ".join(repr(float(x)) for x in c) + "\n")
        for k, blk, i, j, v in body:
            f.write(f"{int(k)} {int(blk)} {int(i)} {int(j)} {float(v)!r}\n")


def raw_from_sdpa_arrays(n, bs, c, body):
    """SDPA (min c'x, sum F_k x_k - F_0 >= 0) -> RawProblem with the MOI wrapper's conventions:
    A[i,1] = F_0, A[i,k+1] = F_k (src/MOI_wrapper.jl:186-196); b = -c (:206); C_lin = -coeff' (:149); d_lin = -F_0[r,r] (:217)."""
    body = np.asarray(body, dtype=np.float64).reshape(-1, 5)
    kk = body[:, 0].astype(np.int64)
    blk = body[:, 1].astype(np.int64)
    ii = body[:, 2].astype(np.int64) - 1
    jj = body[:, 3].astype(np.int64) - 1
    vv = body[:, 4]
    msizes, A = [], []
    lin_blocks, nlin = [], 0
    for b_id, s in enumerate(bs, start=1):
        if s < 0:
            lin_blocks.append((b_id, nlin))
            nlin += -s
    for b_id, s in enumerate(bs, start=1):
        if s < 0:
            continue
        sel = np.nonzero(blk == b_id)[0]
        k, p, q, v = kk[sel], ii[sel], jj[sel], vv[sel]
        od = p != q
        A.append(dict(k=np.concatenate([k, k[od]]), p=np.concatenate([p, q[od]]), q=np.concatenate([q, p[od]]),
                      v=np.concatenate([v, v[od]])))
        msizes.append(int(s))
    if nlin:
        d_lin = np.zeros(nlin)
        R, Cc, V = [], [], []
        for b_id, off in lin_blocks:
            sel = np.nonzero(blk == b_id)[0]
            k, p, q, v = kk[sel], ii[sel], jj[sel], vv[sel]
            if np.any(p != q):
                raise ValueError("off-diagonal entry in a diagonal SDPA block")
            z = k == 0
            np.add.at(d_lin, off + p[z], -v[z])
            R.append(off + p[~z]); Cc.append(k[~z] - 1); V.append(v[~z])
        coeff = sp.csr_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(Cc))), shape=(nlin, n))
        C_lin = (-coeff.T).tocsc()
    else:
        d_lin, C_lin = np.zeros(0), sp.csc_matrix((n, 0))
    return RawProblem(n=n, msizes=msizes, A=A, b=-np.asarray(c, float), b_const=0.0, C_lin=C_lin, d_lin=d_lin)


@dataclass
class MyModel:
    """src/model.jl:34-87."""
    AA: list            # per block scipy csc  n x m^2  (row k = vec(calA_k), math sign)      prep_AA!, src/model.jl:199-229
    B: list             # per block scipy csc  n x m    (datarank = -1)                       prep_B,   src/model.jl:176-197
    C: list             # per block scipy csc  m x m    (= -A[i,1])                           src/model.jl:133
    nzA: np.ndarray
    sigmaA: np.ndarray
    qA: np.ndarray
    b: np.ndarray
    b_const: float
    d_lin: np.ndarray
    C_lin: sp.csc_matrix
    n: int
    msizes: list
    nlin: int
    nlmi: int


def prep_sparse(nz, kappa):
    """src/model.jl:153-174 (nnz-descending stable order; q = number of matrices with nnz > kappa)."""
    order = np.argsort(-nz, kind="stable")
    below = np.nonzero(nz[order] <= kappa)[0]
    q = int(below[0]) if below.size else int(nz.shape[0])
    return order, q


def prep_B(blk, n, m):
    """src/model.jl:176-197: rank-one factors b_k of A[i,k+1] = b_k b_k' (error if not rank one within 5e-6)."""
    k, p, q, v = blk["k"], blk["p"], blk["q"], blk["v"]
    keep = k >= 1
    k, p, q, v = k[keep] - 1, p[keep], q[keep], v[keep]
    order = np.argsort(k, kind="stable")
    k, p, q, v = k[order], p[order], q[order], v[order]
    starts = np.searchsorted(k, np.arange(n + 1))
    rows, cols, vals = [], [], []
    # fast path: single diagonal entry (max-cut style)
    for kk in range(n):
        s, e = starts[kk], starts[kk + 1]
        if e == s:
            continue
        if e - s == 1 and p[s] == q[s]:
            if v[s] < 0:
                raise ValueError("matrix is not rank one positive semidefinite; use datarank = 0")
            rows.append(kk); cols.append(int(p[s])); vals.append(float(np.sqrt(v[s])))
            continue
        idx = np.unique(p[s:e])
        pos = {int(g): t for t, g in enumerate(idx)}
        tmp = np.zeros((idx.size, idx.size))
        for a, bq, val in zip(p[s:e], q[s:e], v[s:e]):
            if int(bq) in pos:
                tmp[pos[int(a)], pos[int(bq)]] += val
        _, vec_ = np.linalg.eigh((tmp + tmp.T) / 2)
        bbb = np.sign(vec_[:, -1]) * np.sqrt(np.abs(np.diag(tmp)))
        err = np.linalg.norm(tmp - np.outer(bbb, bbb))
        if not err <= 5.0e-6:
            raise ValueError(f"Obtained an error of `{err} > 5e-6` when converting matrix into rank `1`, "
                             "use `datarank = 0` to disable the rank-1 conversion.")
        rows += [kk] * idx.size
        cols += [int(g) for g in idx]
        vals += [float(x) for x in bbb]
    return sp.csc_matrix((vals, (rows, cols)), shape=(n, m))


def prepare_model(raw: RawProblem, datarank=0, kappa=8) -> MyModel:
    """src/MOI_wrapper.jl:209-223 + `_prepare_A` (src/model.jl:120-150)."""
    n, nlmi = raw.n, raw.nlmi
    AA, B, C = [], [], []
    nzA = np.zeros((n, nlmi), dtype=np.int64)
    sigmaA = np.zeros((n, nlmi), dtype=np.int64)
    qA = np.zeros((2, nlmi), dtype=np.int64)
    for i in range(nlmi):
        t, m = raw.A[i], int(raw.msizes[i])
        z = t["k"] == 0
        C.append(sp.csc_matrix((-t["v"][z], (t["p"][z], t["q"][z])), shape=(m, m)))
        nk = ~z
        AA.append(sp.csc_matrix((-t["v"][nk], (t["k"][nk] - 1, t["p"][nk] + t["q"][nk] * m)), shape=(n, m * m)))
        if datarank == -1:
            B.append(prep_B(t, n, m))
        nzA[:, i] = np.bincount(t["k"][nk] - 1, minlength=n)
        sigmaA[:, i], q = prep_sparse(nzA[:, i], kappa)
        qA[:, i] = q
    C_lin = raw.C_lin.tocsc() if raw.C_lin is not None else sp.csc_matrix((n, 0))
    return MyModel(AA, B, C, nzA, sigmaA, qA, np.asarray(raw.b, float), float(raw.b_const), np.asarray(raw.d_lin, float),
                   C_lin, n, [int(m) for m in raw.msizes], int(raw.nlin), nlmi)
