"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

SDPA sparse-format (.dat-s) reader + the MOI-equivalent model builder used by the oracle.

What the reference does with an SDPA file (examples/solve_sdpa.jl:14-34 -> MOI.FileFormats.SDPA ->
src/MOI_wrapper.jl:142-232):
  * SDPA problem:   min c'x  s.t.  sum_k F_k x_k - F_0 >= 0   (block diagonal; negative block size = diagonal/LP block)
  * MOI:            PSD block   -> VAF-in-PSDConeTriangle with constants = -F_0
                    diag block  -> VAF-in-Nonnegatives    with constants = -F_0[r,r]
  * copy_to:        A[i,1]   = -constants = F_0      (src/MOI_wrapper.jl:186-189)
                    A[i,k+1] = F_k  (both triangles) (src/MOI_wrapper.jl:190-196)
                    b = -c (MIN sense)               (src/MOI_wrapper.jl:206)
                    C_lin = -coeff'  (n x nlin)      (src/MOI_wrapper.jl:149)
                    d_lin = constants = -F_0[r,r]    (src/MOI_wrapper.jl:217)
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


@dataclass
class RawProblem:
    """Problem exactly as `MOI.copy_to` hands it to `Solvers._prepare_A` / `MyModel` (raw signs).

    A[i] : triplets (k, p, q, v) of the nlmi x (n+1) matrix-of-sparse-matrices `A`, k = 0..n (k=0 is A[i,1]),
           0-based p,q, BOTH triangles stored (as `_add` does, src/MOI_wrapper.jl:178-184).
    """
    n: int
    msizes: list
    A: list                                  # per block: dict(k=int64[], p=int64[], q=int64[], v=float64[])
    b: np.ndarray                            # length n
    b_const: float = 0.0
    C_lin: sp.csc_matrix | None = None       # n x nlin
    d_lin: np.ndarray = field(default_factory=lambda: np.zeros(0))

    @property
    def nlmi(self):
        return len(self.msizes)

    @property
    def nlin(self):
        return int(self.d_lin.shape[0])


_num = re.compile(r"[-+]?(?:\d+\.?\d*(?:[eEdD][-+]?\d+)?|\.\d+(?:[eEdD][-+]?\d+)?)")


def _numbers(line):
    return [float(t.replace("d", "e").replace("D", "e")) for t in _num.findall(line)]


def parse_sdpa(path):
    """Returns (n, blocksizes(signed), c, entries[k, blk, i, j, v]) with 1-based blk/i/j as in the file."""
    with open(path) as f:
        lines = [ln for ln in f if ln.strip() and ln.lstrip()[0] not in '"*']
    n = int(_numbers(lines[0])[0])
    nblocks = int(_numbers(lines[1])[0])
    bs = [int(v) for v in _numbers(lines[2])][:nblocks]
    pos = 3
    c = []
    while len(c) < n:
        c += _numbers(lines[pos])
        pos += 1
    c = np.array(c[:n], dtype=np.float64)
    body = np.array([_numbers(ln)[:5] for ln in lines[pos:] if len(_numbers(ln)) >= 5], dtype=np.float64)
    if body.size == 0:
        body = np.zeros((0, 5))
    return n, bs, c, body


def raw_from_sdpa(path):
    n, bs, c, body = parse_sdpa(path)
    return raw_from_sdpa_arrays(n, bs, c, body)


def raw_from_sdpa_arrays(n, bs, c, body):
    k = body[:, 0].astype(np.int64)
    blk = body[:, 1].astype(np.int64)
    ii = body[:, 2].astype(np.int64) - 1
    jj = body[:, 3].astype(np.int64) - 1
    v = body[:, 4]
    msizes, A = [], []
    lin_off, nlin = {}, 0
    for bidx, s in enumerate(bs, start=1):
        if s < 0:
            lin_off[bidx] = nlin
            nlin += -s
    for bidx, s in enumerate(bs, start=1):
        if s < 0:
            continue
        sel = blk == bidx
        kk, p, q, vv = k[sel], ii[sel], jj[sel], v[sel]
        off = p != q
        # both triangles, src/MOI_wrapper.jl:178-184
        A.append(dict(k=np.concatenate([kk, kk[off]]), p=np.concatenate([p, q[off]]),
                      q=np.concatenate([q, p[off]]), v=np.concatenate([vv, vv[off]])))
        msizes.append(int(s))
    # LP rows: coefficient of x_k in row r is F_k[r,r]; constant is -F_0[r,r]
    if nlin > 0:
        rows, cols, vals = [], [], []
        d_lin = np.zeros(nlin)
        for bidx, off in lin_off.items():
            sel = blk == bidx
            kk, p, q, vv = k[sel], ii[sel], jj[sel], v[sel]
            assert np.all(p == q), "off-diagonal entry in a diagonal SDPA block"
            z = kk == 0
            np.add.at(d_lin, off + p[z], -vv[z])
            rows.append(off + p[~z]); cols.append(kk[~z] - 1); vals.append(vv[~z])
        coeff = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nlin, n))
        C_lin = (-coeff.T).tocsc()
    else:
        C_lin, d_lin = sp.csc_matrix((n, 0)), np.zeros(0)
    return RawProblem(n=n, msizes=msizes, A=A, b=-c, b_const=0.0, C_lin=C_lin, d_lin=d_lin)
