"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path (loraine.jl_b200/).

CPU (NumPy/SciPy) restatement of the Loraine.jl v0.2.5 interior-point hot path, function by function.
Every function cites the reference lines it restates (paths relative to /root/reference).

PARITY PINNING.  The reference is pure Julia and cannot be executed in the build container (no Julia, no
network), and it ships no unit tests / golden vectors for any hot-path function.  What pins this oracle:
  * the reference's own end-to-end known answers (tests/test_oracle_golden.py):
      theta1 objective 23 rtol 1e-6 (examples/solve_sdpa.jl:61), ex_corr -0.9779977649 / 0.8719210472
      (examples/ex_corr.jl:30-31), ex_dist 4/3 + Q matrix (examples/ex_dist.jl:29-40), 4-node max-cut partition
      (examples/ex_maxcut.jl:46-47), k.jl LP optimum 4 / x=2 (examples/k.jl:29-38);
  * SDPLIB optima of the bundled .dat-s files (control1 17.78463, maxG11 629.1648, thetaG11 400.00).
Intermediate quantities (H, W, dely, CG / IP iteration counts) are NOT pinned by any reference fixture:
"parity unpinned" at function granularity; the "+-1 IP iteration vs the reference's own CPU run" criterion
stays open until a Julia run exists.

Third-party arithmetic that is not under /root/reference (Project.toml:6-31, no Manifest => versions unpinned):
  * ConjugateGradients.jl 0.1 `cg`   -> `cg` below restates its published PCG recurrence / exit codes.
  * FameSVD.jl 0.1 `fsvd`            -> `fsvd` below: SVD through the symmetric eigenproblem of A'A.
  * Julia LinearAlgebra (LAPACK)     -> scipy.linalg (LAPACK) cholesky / eigh / solve_triangular.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------------------------
# options  (src/Solvers.jl:169-185)
# --------------------------------------------------------------------------------------------
DEFAULT_OPTIONS = dict(kit=0, tol_cg=1.0e-2, tol_cg_up=0.5, tol_cg_min=1.0e-7, eDIMACS=1.0e-7, preconditioner=1,
                       erank=1, aamat=1, fig_ev=0, verb=1, datarank=0, initpoint=0, timing=1, maxit=100,
                       datasparsity=8)


# --------------------------------------------------------------------------------------------
# kron_etc.jl
# --------------------------------------------------------------------------------------------
def vec(M):
    return np.asarray(M).reshape(-1, order="F")


def mat(v):
    """src/kron_etc.jl:13-18  reshape (column-major) + symmetrise."""
    n = math.isqrt(v.shape[0])
    A = np.asarray(v).reshape(n, n, order="F")
    return (A + A.T) / 2


def my_kron(A, B, C):
    """src/kron_etc.jl:4-11  vec(B*C*A')."""
    return vec(B @ (C @ A.T))


def btrace(nlmi, X, S):
    """src/kron_etc.jl:21-28."""
    t = 0.0
    for i in range(nlmi):
        Si = S[i]
        t += float(X[i].multiply(Si).sum()) if sp.issparse(X[i]) else float(np.sum(X[i] * (Si.toarray() if sp.issparse(Si) else Si)))
    return t


# --------------------------------------------------------------------------------------------
# model.jl
# --------------------------------------------------------------------------------------------
@dataclass
class MyModel:
    """src/model.jl:34-87 (fields) built by _prepare_A (src/model.jl:120-150)."""
    raw: object
    AA: list            # per block csr n x m^2, row k = vec(calA_k) = -vec(A[i,k+1])   (src/model.jl:199-229,138)
    B: list             # per block csr n x m (datarank == -1)                       (src/model.jl:176-197)
    C: list             # per block csc m x m, C_i = -A[i,1]                         (src/model.jl:133)
    nzA: np.ndarray     # n x nlmi
    sigmaA: np.ndarray  # n x nlmi, 0-based permutation, nnz-descending, stable     (src/model.jl:160)
    qA: np.ndarray      # 2 x nlmi
    b: np.ndarray
    b_const: float
    d_lin: np.ndarray
    C_lin: sp.csr_matrix  # n x nlin
    n: int
    msizes: list
    nlin: int
    nlmi: int
    _Acache: dict = field(default_factory=dict)

    def A(self, i, k):
        """A[i,k] (k = 0..n, raw sign) as a csc m x m matrix -- built lazily, used only by the as-written loops."""
        key = (i, k)
        if key not in self._Acache:
            t = self.raw.A[i]
            m = self.msizes[i]
            if "_order" not in t:
                o = np.argsort(t["k"], kind="stable")
                t["_order"] = o
                t["_ptr"] = np.searchsorted(t["k"][o], np.arange(self.n + 2))
            o, ptr = t["_order"], t["_ptr"]
            s = o[ptr[k]:ptr[k + 1]]
            self._Acache[key] = sp.csc_matrix((t["v"][s], (t["p"][s], t["q"][s])), shape=(m, m))
        return self._Acache[key]


def prep_sparse(nz, kappa):
    """src/model.jl:153-174."""
    n = nz.shape[0]
    sigma = np.argsort(-nz, kind="stable")
    sisi = nz[sigma]
    q = n
    for j in range(n):
        if sisi[j] <= kappa:
            q = j
            break
    return sigma, q


def prep_B(model_raw_block, n, m):
    """src/model.jl:176-197: A[i,k+1] = b_k b_k' (rank one, raw sign); row k of B = b_k."""
    t = model_raw_block
    rows, cols, vals = [], [], []
    o = np.argsort(t["k"], kind="stable")
    ptr = np.searchsorted(t["k"][o], np.arange(n + 2))
    for k in range(1, n + 1):
        s = o[ptr[k]:ptr[k + 1]]
        if s.size == 0:
            continue
        bidx = np.unique(t["p"][s])
        loc = {int(g): l for l, g in enumerate(bidx)}
        tmp = np.zeros((bidx.size, bidx.size))
        for p, q, v in zip(t["p"][s], t["q"][s], t["v"][s]):
            if int(q) in loc:
                tmp[loc[int(p)], loc[int(q)]] += v
        _, vtmp = np.linalg.eigh((tmp + tmp.T) / 2)
        bbb = np.sign(vtmp[:, -1]) * np.sqrt(np.diag(tmp))
        err = np.linalg.norm(tmp - np.outer(bbb, bbb))
        if not err <= 5.0e-6:
            raise ValueError(f"Obtained an error of `{err} > 5e-6` when converting matrix into rank `1`, "
                             "use `datarank = 0` to disable the rank-1 conversion.")
        rows += [k - 1] * bidx.size
        cols += list(bidx)
        vals += list(bbb)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, m))


def prepare_model(raw, datarank=0, kappa=8):
    """src/MOI_wrapper.jl:209-223 + src/model.jl:120-150 (_prepare_A)."""
    n, nlmi = raw.n, raw.nlmi
    AA, B, C = [], [], []
    nzA = np.zeros((n, nlmi), dtype=np.int64)
    sigmaA = np.zeros((n, nlmi), dtype=np.int64)
    qA = np.zeros((2, nlmi), dtype=np.int64)
    for i in range(nlmi):
        t, m = raw.A[i], raw.msizes[i]
        z = t["k"] == 0
        C.append(sp.csc_matrix((-t["v"][z], (t["p"][z], t["q"][z])), shape=(m, m)))
        nzk = ~z
        AA.append(sp.csr_matrix((-t["v"][nzk], (t["k"][nzk] - 1, t["p"][nzk] + t["q"][nzk] * m)), shape=(n, m * m)))
        if datarank == -1:
            B.append(prep_B(t, n, m))
        nzA[:, i] = np.bincount(t["k"][nzk] - 1, minlength=n)
        sigmaA[:, i], q = prep_sparse(nzA[:, i], kappa)
        qA[0, i] = qA[1, i] = q
    C_lin = raw.C_lin.tocsr() if raw.C_lin is not None else sp.csr_matrix((n, 0))
    return MyModel(raw, AA, B, C, nzA, sigmaA, qA, np.asarray(raw.b, float), float(raw.b_const),
                   np.asarray(raw.d_lin, float), C_lin, n, list(raw.msizes), int(raw.nlin), nlmi)


# --------------------------------------------------------------------------------------------
# solver state (src/Solvers.jl:18-147) -- plain attribute bag
# --------------------------------------------------------------------------------------------
class MySolver:
    pass


class Halpha:
    pass


def load(model, options):
    """src/Solvers.jl:187-302 (incl. the parameter range checks :263-291)."""
    o = dict(DEFAULT_OPTIONS)
    o.update(options)
    s = MySolver()
    for k in ("kit", "preconditioner", "erank", "aamat", "verb", "datarank", "initpoint", "maxit", "datasparsity"):
        setattr(s, k, int(o[k]))
    for k in ("tol_cg", "tol_cg_up", "tol_cg_min", "eDIMACS"):
        setattr(s, k, float(o[k]))
    s.model = model
    s.cg_iter_tot = 0
    if s.kit < 0 or s.kit > 1:
        s.kit = 0
    if s.tol_cg < s.tol_cg_min and s.kit == 1:
        s.tol_cg = s.tol_cg_min
    if s.tol_cg_min > s.eDIMACS and s.kit == 1:
        s.tol_cg_min = s.eDIMACS
    if s.kit == 1 and (s.preconditioner < 0 or s.preconditioner > 4):
        s.preconditioner = 1
    if s.erank < 0:
        s.erank = 1
    if s.datarank < -1:
        s.datarank = 0
    if s.initpoint < 0 or s.initpoint > 1:
        s.initpoint = 1
    s.trace = []
    s.hooks = {}
    s.phase_time = {}
    return s, Halpha()


def setup_solver(s, halpha):
    """src/Solvers.jl:363-446."""
    md = s.model
    z = lambda: [np.zeros((m, m)) for m in md.msizes]
    s.X, s.S, s.delX, s.delS, s.G, s.Gi, s.W, s.Si, s.Rd, s.Xn, s.Sn, s.RNT = (z() for _ in range(12))
    s.D = [np.zeros(m) for m in md.msizes]
    s.DDsi = [np.zeros(m) for m in md.msizes]
    s.alpha = np.zeros(md.nlmi)
    s.beta = np.zeros(md.nlmi)
    s.regcount = 0
    halpha.Umat = [np.zeros((m, s.erank)) for m in md.msizes]
    halpha.Z = [np.zeros((m, m)) for m in md.msizes]
    if s.kit == 1:
        if md.nlmi == 0:
            s.kit = 0
        elif s.erank >= max(md.msizes) - 1:
            s.kit = 0
    if len(md.B) > 0:
        for i in range(md.nlmi):
            if md.B[i].nnz == 0:
                s.datarank = 0
    s.X_lin = s.S_lin = s.Si_lin = s.S_lin_inv = np.zeros(0)
    s.delX_lin = s.delS_lin = s.Xn_lin = s.Sn_lin = s.RNT_lin = s.Rd_lin = np.zeros(0)
    s.alpha_lin = s.beta_lin = 1.0
    s.cholBBBB = None
    s.chol_is_factor_object = False


# --------------------------------------------------------------------------------------------
# initial_point.jl
# --------------------------------------------------------------------------------------------
def initial_point(s):
    """src/initial_point.jl:1-81."""
    md = s.model
    n = md.b.shape[0]
    s.y = np.zeros(n)
    b2 = 1 + np.abs(md.b)
    f = 0.0
    for i in range(md.nlmi):
        m = md.msizes[i]
        if s.initpoint == 0:
            Eps = 1.0
        else:
            f = np.linalg.norm(b2) / (1 + spla.norm(md.AA[i]))
            Eps = math.sqrt(m) * max(1.0, math.sqrt(m) * f)
        s.X[i] = Eps * np.eye(m)
        if s.initpoint == 0:
            Eta = float(md.n)
        else:
            mf = max(f, spla.norm(md.C[i]))
            mf = (1 + mf) / math.sqrt(m)
            Eta = math.sqrt(m) * max(1.0, mf)
        s.S[i] = Eta * np.eye(m)
    if md.nlin > 0:
        dd = md.d_lin.shape[0]
        rown = np.sqrt(np.asarray(md.C_lin.multiply(md.C_lin).sum(axis=1)).ravel())
        if s.initpoint == 0:
            Epss = 1.0
        else:
            p = b2 / (1 + rown)
            Epss = max(1.0, float(p.max()))
        s.X_lin = Epss * np.ones(dd)
        if s.initpoint == 0:
            Etaa = 1.0
        else:
            mf = max(float(rown.max()), float(np.linalg.norm(md.d_lin)))
            mf = mf / math.sqrt(dd)
            Etaa = max(1.0, mf)
        s.S_lin = Etaa * np.ones(dd)
        s.S_lin_inv = 1.0 / s.S_lin
    s.sigma = 3.0
    s.tau = 0.95
    s.expon = 3.0
    s.DIMACS_error = 1.0
    s.iter = 0
    s.status = 0


# --------------------------------------------------------------------------------------------
# Solvers.jl: find_mu, check_convergence
# --------------------------------------------------------------------------------------------
def find_mu(s):
    """src/Solvers.jl:480-494."""
    md = s.model
    tr = 0.0
    for i in range(md.nlmi):
        tr += float(np.sum(s.X[i] * s.S[i]))
    if md.nlin > 0:
        tr += float(s.X_lin @ s.S_lin)
    s.mu = tr / (sum(md.msizes) + md.nlin)
    return s.mu


def eigmin(M):
    return float(sla.eigh(M, eigvals_only=True, subset_by_index=[0, 0])[0])


def check_convergence(s):
    """src/Solvers.jl:496-568 (uses the residuals Rp/Rd left by the predictor, i.e. lagged)."""
    md = s.model
    nb = np.linalg.norm(md.b)
    by = float(md.b @ s.y)
    s.err1 = float(np.linalg.norm(s.Rp) / (1 + nb))
    s.err2 = s.err3 = s.err4 = s.err5 = s.err6 = 0.0
    trCX = 0.0
    for i in range(md.nlmi):
        nC = spla.norm(md.C[i])
        CX = float(md.C[i].multiply(s.X[i]).sum())
        trCX += CX
        s.err2 += max(0.0, -eigmin(s.X[i]) / (1 + nb))
        s.err3 += float(np.linalg.norm(s.Rd[i]) / (1 + nC))
        s.err4 += max(0.0, -eigmin(s.S[i]) / (1 + nC))
        s.err6 += float(np.sum(s.S[i] * s.X[i])) / (1 + abs(CX) + abs(by))
    s.err5 = (trCX - by) / (1 + abs(trCX) + abs(by))
    if md.nlin > 0:
        nd = np.linalg.norm(md.d_lin)
        dx = float(md.d_lin @ s.X_lin)
        s.err2 += max(0.0, -float(s.X_lin.min()) / (1 + nb))
        s.err3 += float(np.linalg.norm(s.Rd_lin) / (1 + nd))
        s.err4 += max(0.0, -float(s.S_lin.min()) / (1 + nd))
        s.err5 = (trCX + dx - by) / (1 + abs(trCX) + abs(by))
        s.err6 += float(s.S_lin @ s.X_lin) / (1 + abs(dx) + abs(by))
    else:
        dx = 0.0
    if md.nlmi > 0:
        D = s.err1 + s.err2 + s.err3 + s.err4 + abs(s.err5) + s.err6
    else:
        D = s.err2 + s.err3 + s.err4 + abs(s.err5) + s.err6
    s.DIMACS_error = D
    s.primal_obj = -by + md.b_const          # printed "obj", src/Solvers.jl:530
    s.dual_obj = -trCX - dx                  # src/Solvers.jl:549
    if D < s.eDIMACS:
        s.status = 1
    if D > 1e55:
        s.status = 2
    elif abs(by) > 1e55:
        s.status = 3
    return D


# --------------------------------------------------------------------------------------------
# prepare_W.jl
# --------------------------------------------------------------------------------------------
def fsvd(A):
    """FameSVD.fsvd (un-vendored, compat 0.1): SVD of a square matrix through the symmetric eigenproblem of A'A,
    singular values descending; U = A V / s.  (src/prepare_W.jl:42 is the only call site; only D and V are used.)"""
    w, V = sla.eigh(A.T @ A)
    o = np.argsort(-w, kind="stable")
    w, V = w[o], V[:, o]
    sv = np.sqrt(np.maximum(w, 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        U = (A @ V) / sv
    return U, sv, V


def try_cholesky(s, Xl, i, status_on_fail=4):
    """src/prepare_W.jl:5-26: lower Cholesky factor; on failure add 1e-5*I (in place!) up to 1000 times."""
    try:
        return np.linalg.cholesky(Xl[i])
    except np.linalg.LinAlgError:
        icount = 0
        while True:
            try:
                return np.linalg.cholesky(Xl[i])
            except np.linalg.LinAlgError:
                Xl[i] = Xl[i] + 1e-5 * np.eye(Xl[i].shape[0])
                icount += 1
                if icount > 1000:
                    s.status = status_on_fail
                    return np.eye(Xl[i].shape[0])


def prepare_W(s):
    """src/prepare_W.jl:28-94."""
    md = s.model
    for i in range(md.nlmi):
        m = md.msizes[i]
        L = try_cholesky(s, s.X, i)
        LS = try_cholesky(s, s.S, i)
        CC = LS.T @ L
        _, Dtmp, V = fsvd(CC)
        s.D[i] = Dtmp.copy()
        Di2 = 1.0 / np.sqrt(Dtmp)
        s.G[i] = (L @ V) * Di2[None, :]
        s.Gi[i] = np.linalg.inv(s.G[i])
        s.W[i] = s.G[i] @ s.G[i].T
        s.Si[i] = sla.solve_triangular(LS.T, sla.solve_triangular(LS, np.eye(m), lower=True), lower=False)
        DD = s.G[i].T @ s.S[i] @ s.G[i]
        DD = (DD + DD.T) / 2
        s.DDsi[i] = 1.0 / np.sqrt(np.diag(DD))
    s.Si_lin = 1.0 / s.S_lin if md.nlin > 0 else np.zeros(0)


# --------------------------------------------------------------------------------------------
# makeBBBB.jl
# --------------------------------------------------------------------------------------------
def makeBBBB_rank1(n, nlmi, B, G):
    """src/makeBBBB.jl:1-20:  H = sum_i ((B_i G_i)(B_i G_i)').^2."""
    H = np.zeros((n, n))
    for i in range(nlmi):
        BG = B[i] @ G[i]
        tmp = BG @ BG.T
        H += tmp ** 2
    return H


def _dot(A, Bm, W):
    """src/makeBBBB.jl:39-64  <A W, W B> for sparse symmetric csc A, B -- literal triple loop."""
    m = W.shape[0]
    res = 0.0
    Ap, Ai, Ax = A.indptr, A.indices, A.data
    Bp, Bi, Bx = Bm.indptr, Bm.indices, Bm.data
    for i in range(m):
        if Ap[i + 1] > Ap[i]:
            for j in range(m):
                if Bp[j + 1] > Bp[j]:
                    AW = 0.0
                    for k in range(Ap[i], Ap[i + 1]):
                        AW += Ax[k] * W[Ai[k], j]
                    WB = 0.0
                    for k in range(Bp[j], Bp[j + 1]):
                        WB += W[i, Bi[k]] * Bx[k]
                    res += AW * WB
    return res


def makeBBBBsi_aswritten(md, ilmi, W):
    """src/makeBBBB.jl:67-218 literally (F1 for the first qA[1] matrices in nnz order, F3 otherwise; F2 is dead code).
    Only the lower triangle is guaranteed to be filled by the F3 branch (:178-182, :202-206).  Small sizes only."""
    n = md.n
    H = np.zeros((n, n))
    sig = md.sigmaA[:, ilmi]
    AAi = md.AA[ilmi]
    for ii in range(n):
        i = sig[ii]
        Ai = md.A(ilmi, i + 1)
        if Ai.nnz > 0:
            if ii < md.qA[0, ilmi]:
                tmp = (W @ Ai.toarray()) @ W                       # :88,:92
                tmp2 = AAi @ vec(tmp)                                # :95
                indi = sig[ii:]
                H[indi, i] = -tmp2[indi]                             # :100
                H[i, indi] = -tmp2[indi]                             # :101
            else:
                if Ai.nnz > 1:
                    for jj in range(ii, n):
                        j = sig[jj]
                        Aj = md.A(ilmi, j + 1)
                        if Aj.nnz > 0:
                            ttt = _dot(Ai, Aj, W)
                            if i >= j:
                                H[i, j] = ttt
                            else:
                                H[j, i] = ttt
                else:
                    a = int(Ai.indices[0])
                    vi = float(Ai.data[0])
                    for jj in range(ii, n):
                        j = sig[jj]
                        Aj = md.A(ilmi, j + 1)
                        if Aj.nnz > 0:
                            c = int(Aj.indices[0])
                            vj = float(Aj.data[0])
                            ttt = vi * W[a, c] * W[a, c] * vj
                            if i >= j:
                                H[i, j] = ttt
                            else:
                                H[j, i] = ttt
    return H


def makeBBBBsi_entries(md, ilmi, W, chunk=2048):
    """Vectorised evaluation of the SAME matrix H_i[j,k] = tr(calA_j W calA_k W) (the quantity every branch of
    src/makeBBBB.jl:67-218 computes) through the F3 identity (:172) summed over stored entries:
        H_i[j,k] = sum_{(a,b) in j} sum_{(p,q) in k} v_j[a,b] v_k[p,q] W[a,p] W[b,q].
    Full symmetric matrix; used for medium-size oracle runs where the literal loops would take hours."""
    n, m = md.n, md.msizes[ilmi]
    AAi = md.AA[ilmi].tocoo()
    r, c, v = AAi.row, AAi.col, AAi.data
    a, b = c % m, c // m
    P = sp.csr_matrix((v, (r, np.arange(r.size))), shape=(n, r.size))
    H = np.zeros((n, n))
    for s0 in range(0, r.size, chunk):
        sl = slice(s0, min(s0 + chunk, r.size))
        M = W[a[sl]][:, a] * W[b[sl]][:, b]           # |E| x nnz
        H += P[:, sl] @ (P @ M.T).T
    return (H + H.T) / 2


def makeBBBBs(md, W, aswritten=False):
    """src/makeBBBB.jl:24-36."""
    H = np.zeros((md.n, md.n))
    for i in range(md.nlmi):
        H += makeBBBBsi_aswritten(md, i, W[i]) if aswritten else makeBBBBsi_entries(md, i, W[i])
    return H


def makeRHS(nlmi, AA, W, S, Rp, Rd):
    """src/makeBBBB.jl:221-228."""
    h = Rp
    for i in range(nlmi):
        h = h + AA[i] @ vec(W[i] @ (Rd[i] + S[i]) @ W[i])
    return h


def lp_schur(md, x_over_s):
    """src/predictor_corrector.jl:36-38:  C_lin * spdiagm(x./s) * C_lin'."""
    return (md.C_lin @ sp.diags(x_over_s) @ md.C_lin.T).toarray()


# --------------------------------------------------------------------------------------------
# ConjugateGradients.jl `cg` (un-vendored dependency; published recurrence restated)
# --------------------------------------------------------------------------------------------
def cg(A, b, tol=1e-6, maxIter=100, precon=None):
    """x0 = 0; returns (x, exit_code, num_iters).  Call sites: src/predictor_corrector.jl:134,235."""
    n = b.shape[0]
    x = np.zeros(n)
    if np.linalg.norm(b) == 0.0:
        return x, 1, 0
    r = b - A(x)
    res0 = np.linalg.norm(r)
    if res0 <= tol:
        return x, 2, 0
    z = precon(r) if precon is not None else r.copy()
    p = z.copy()
    for it in range(1, maxIter + 1):
        Ap = A(p)
        gamma = float(r @ z)
        alpha = gamma / float(p @ Ap)
        if alpha == math.inf or alpha < 0:
            return x, -13, it
        x = x + alpha * p
        r = r - alpha * Ap
        if np.linalg.norm(r) / res0 <= tol:
            return x, 30, it
        z = precon(r) if precon is not None else r.copy()
        beta = float(z @ r) / gamma
        p = z + beta * p
    return x, -2, maxIter


# --------------------------------------------------------------------------------------------
# Solvers.jl:570-904  CG operator + preconditioners
# --------------------------------------------------------------------------------------------
def MyA(s):
    """src/Solvers.jl:572-614."""
    md = s.model

    def apply(x):
        ax1 = np.zeros(md.n)
        for i in range(md.nlmi):
            ax = md.AA[i].T @ x
            waxw = s.W[i] @ mat(ax) @ s.W[i]
            ax1 += md.AA[i] @ vec(waxw)
        if md.nlin > 0:
            ax1 += md.C_lin @ ((s.X_lin * s.S_lin_inv) * (md.C_lin.T @ x))
        return ax1
    return apply


def _tau(lambda_s, aamat):
    if aamat == 0:
        return 1.0 * float(lambda_s.min())
    return (float(lambda_s.min()) + float(lambda_s.mean())) / 2 - 1.0e-14


def Prec_for_CG_beta(s, halpha):
    """src/Solvers.jl:624-663."""
    md = s.model
    d = np.zeros(md.n)
    for i in range(md.nlmi):
        m = md.msizes[i]
        lam = sla.eigh(s.W[i], eigvals_only=True)
        ttau = _tau(lam[: m - s.erank], s.aamat)
        if s.aamat < 3:
            d += ttau ** 2
    if md.nlmi > 0 and md.nlin > 0:
        d += np.asarray((md.C_lin.multiply(md.C_lin)) @ (s.X_lin * s.S_lin_inv)).ravel()
    halpha.AAAATtau = d


def MyM_beta(halpha):
    """src/Solvers.jl:665-672."""
    return lambda x: x / halpha.AAAATtau


def Prec_for_CG_tilS_prep(s, halpha):
    """src/Solvers.jl:674-809 (+ prec_alpha_S! :819-864 for erank == 1)."""
    md = s.model
    nvar, k = md.n, s.erank
    halpha.Z = []
    Dmat = sp.csc_matrix((nvar, nvar))
    for i in range(md.nlmi):
        m = md.msizes[i]
        lam, vect = sla.eigh(s.W[i])
        vect_l, lambda_l = vect[:, m - k:], lam[m - k:]
        vect_s, lambda_s = vect[:, : m - k], lam[: m - k]
        ttau = _tau(lambda_s, s.aamat)
        halpha.Umat[i] = vect_l * np.sqrt(lambda_l - ttau)[None, :]
        W0 = (vect_s * lambda_s[None, :]) @ vect_s.T + ttau * (vect_l @ vect_l.T)
        W0 = (W0 + W0.T) / 2
        halpha.Z.append(np.linalg.cholesky(2 * W0 + halpha.Umat[i] @ halpha.Umat[i].T))
        if s.aamat < 3:
            Dmat = Dmat + (ttau ** 2) * sp.identity(nvar, format="csc")
    if md.nlin > 0:
        Dmat = Dmat + (md.C_lin @ sp.diags(s.X_lin * s.S_lin_inv) @ md.C_lin.T).tocsc()
    halpha.AAAATtau = Dmat.tocsc()
    halpha.diag_only = md.nlin == 0
    halpha.Ddiag = Dmat.diagonal()
    halpha.Dsolve = (lambda x: x / halpha.Ddiag) if halpha.diag_only else spla.factorized(halpha.AAAATtau)
    if k > 1:                                                       # :752-768 slow formula
        cols = []
        for i in range(md.nlmi):
            cols.append(md.AA[i] @ np.kron(halpha.Umat[i], halpha.Z[i]))
        t = np.hstack(cols)
        Dt = np.column_stack([halpha.Dsolve(t[:, j]) for j in range(t.shape[1])])
        S = t.T @ Dt
    else:                                                           # :770,:800 fast formula, diag(D) only
        dsq = np.sqrt(1.0 / halpha.Ddiag)
        cols = []
        for i in range(md.nlmi):
            m = md.msizes[i]
            AAs = sp.diags(dsq) @ md.AA[i]
            coo = AAs.tocoo()
            qq, pp = coo.col // m, coo.col % m
            AU = sp.csr_matrix((coo.data * halpha.Umat[i][qq, 0], (coo.row, pp)), shape=(nvar, m))
            cols.append(AU @ halpha.Z[i])
        t = np.hstack(cols)
        S = t.T @ t
    S = (S + S.T) / 2 + np.eye(S.shape[0])
    halpha.cholS = sla.cho_factor(S, lower=True)


def MyM(s, halpha):
    """src/Solvers.jl:866-904."""
    md = s.model

    def apply(x):
        v = halpha.Dsolve(x)
        y33 = []
        for i in range(md.nlmi):
            y22 = md.AA[i].T @ v
            y33.append(vec(halpha.Z[i].T @ mat(y22) @ halpha.Umat[i]))
        y33 = sla.cho_solve(halpha.cholS, np.concatenate(y33))
        yy2 = np.zeros(md.n)
        ii = 0
        for i in range(md.nlmi):
            m, k = halpha.Umat[i].shape
            yy = np.zeros(m * m)
            for r in range(k):
                xx = halpha.Z[i] @ y33[ii:ii + m]
                yy += np.kron(halpha.Umat[i][:, r], xx)
                ii += m
            yy2 += md.AA[i] @ yy
        return v - halpha.Dsolve(yy2)
    return apply


# --------------------------------------------------------------------------------------------
# predictor_corrector.jl
# --------------------------------------------------------------------------------------------
class PosDefException(Exception):
    pass


def _chol_lower(H):
    try:
        return np.linalg.cholesky(H)
    except np.linalg.LinAlgError as e:
        raise PosDefException(str(e))


def _chol_lower_inplace(H):
    """LAPACK dpotrf on the lower triangle of a Fortran-ordered H, in place (no n x n temporaries; the strict upper triangle
    is left as it was and never read by the triangular solves) -- same factor as _chol_lower."""
    L, info = sla.lapack.dpotrf(H, lower=1, overwrite_a=1, clean=0)
    if info != 0:
        raise PosDefException(f"dpotrf info = {info}")
    return L


def _assemble_lean(s):
    """kit = 0, datarank != -1 at sizes where the NumPy assembly (makeBBBBsi_entries) and its n x n temporaries take minutes:
    the lower triangle of the same H = sum_i tr(calA_j W_i calA_k W_i) (+ LP term) from the plain-C restatement
    oracle/schur_pairs.c, Fortran-ordered, nothing mirrored (`Hermitian(BBBB, :L)` only reads the lower triangle)."""
    from . import c_oracle
    md = s.model
    H = np.zeros((md.n, md.n), order="F")
    for i in range(md.nlmi):
        c_oracle.schur_pairs_lower(md.AA[i], md.msizes[i], s.W[i], H=H, accumulate=(i > 0), nthreads=getattr(s, "threads", 0))
    if md.nlin > 0:
        H += np.tril(lp_schur(md, s.X_lin * s.S_lin_inv))
    return H


def _timed(s, name, t0):
    s.phase_time[name] = s.phase_time.get(name, 0.0) + (time.perf_counter() - t0)


def _solve_direct(s, h):
    """src/predictor_corrector.jl:89-90 / :199  `cholBBBB' \\ (cholBBBB \\ h)`.
    Normal path: cholBBBB is the LowerTriangular factor L.  After a regularised retry (:85) it is a `Cholesky`
    OBJECT whose adjoint is itself, so the same expression evaluates H^-1 H^-1 h (reference quirk, kept)."""
    if s.chol_is_factor_object:
        x = sla.cho_solve((s.cholBBBB, True), h)
        return sla.cho_solve((s.cholBBBB, True), x)
    x = sla.solve_triangular(s.cholBBBB, h, lower=True, check_finite=False)
    return sla.solve_triangular(s.cholBBBB, x, lower=True, trans="T", check_finite=False)


def predictor(s, halpha):
    """src/predictor_corrector.jl:5-146."""
    md = s.model
    s.predict = True
    t0 = time.perf_counter()
    s.Rp = md.b.copy()
    for i in range(md.nlmi):
        s.Rp = s.Rp - md.AA[i] @ vec(s.X[i])
        s.Rd[i] = md.C[i].toarray() - s.S[i] - mat(md.AA[i].T @ s.y)
    if md.nlin > 0:
        s.Rp = s.Rp - md.C_lin @ s.X_lin
        s.Rd_lin = md.d_lin - s.S_lin - md.C_lin.T @ s.y
    _timed(s, "residuals", t0)

    lean = bool(getattr(s, "lean", False)) and s.kit == 0 and s.datarank != -1 and md.nlmi > 0
    if lean:
        t0 = time.perf_counter()
        BBBB = _assemble_lean(s)
        _timed(s, "schur_assemble", t0)
    elif s.kit == 0:
        t0 = time.perf_counter()
        if md.nlmi > 0:
            if s.datarank == -1:
                BBBB = makeBBBB_rank1(md.n, md.nlmi, md.B, s.G)
            else:
                BBBB = makeBBBBs(md, s.W, aswritten=getattr(s, "aswritten", False))
        else:
            BBBB = np.zeros((md.n, md.n))
        if md.nlin > 0:
            BBBB = BBBB + lp_schur(md, s.X_lin * s.S_lin_inv)
        BBBB = np.tril(BBBB) + np.tril(BBBB, -1).T            # Hermitian(BBBB, :L)  (:39)
        _timed(s, "schur_assemble", t0)
        if "H" in s.hooks:
            s.hooks["H"](s, BBBB)

    t0 = time.perf_counter()
    if md.nlmi > 0:
        h = makeRHS(md.nlmi, md.AA, s.W, s.S, s.Rp, s.Rd)
    else:
        h = s.Rp.copy()
    if md.nlin > 0:
        h = h + md.C_lin @ ((s.X_lin * s.Si_lin) * s.Rd_lin + s.X_lin)
    _timed(s, "rhs", t0)

    if s.kit == 0:
        t0 = time.perf_counter()
        try:
            s.cholBBBB = _chol_lower_inplace(BBBB) if lean else _chol_lower(BBBB)
            s.chol_is_factor_object = False
        except PosDefException:
            if lean:                                              # the in-place attempt destroyed H: form it again
                BBBB = _assemble_lean(s)
                BBBB = np.tril(BBBB) + np.tril(BBBB, -1).T
            icount = 0
            s.regcount += 1
            if s.regcount > 5:
                s.cholBBBB = np.eye(md.n)
                s.status = 3
                return
            while True:
                try:
                    _chol_lower(BBBB)
                    break
                except PosDefException:
                    BBBB = BBBB + 1e-4 * np.eye(md.n)
                    icount += 1
                    if icount > 1000:
                        s.cholBBBB = np.eye(md.n)
                        s.status = 3
                        return
            s.cholBBBB = _chol_lower(BBBB)
            s.chol_is_factor_object = True
        _timed(s, "schur_factor", t0)
        t0 = time.perf_counter()
        s.dely = _solve_direct(s, h)
        _timed(s, "schur_solve", t0)
    else:
        t0 = time.perf_counter()
        A = MyA(s)
        if s.preconditioner == 0:
            M = None
        elif s.preconditioner == 1:
            Prec_for_CG_tilS_prep(s, halpha)
            M = MyM(s, halpha)
        elif s.preconditioner in (2, 4):
            Prec_for_CG_beta(s, halpha)
            M = MyM_beta(halpha)
        else:
            raise ValueError("preconditioner 3 is undefined in the reference (src/predictor_corrector.jl:120-128)")
        _timed(s, "prec_prepare", t0)
        t0 = time.perf_counter()
        s.dely, exit_code, num_iters = cg(A, h, tol=s.tol_cg, maxIter=10000, precon=M)
        _timed(s, "cg", t0)
        s.cg_iter_pre += num_iters
        s.cg_iter_tot += num_iters
    if "dely_pred" in s.hooks:
        s.hooks["dely_pred"](s, h, s.dely)
    t0 = time.perf_counter()
    find_step(s)
    _timed(s, "find_step", t0)


def sigma_update(s):
    """src/predictor_corrector.jl:148-179."""
    md = s.model
    step_pred = min(min(list(s.alpha) + [s.alpha_lin]), min(list(s.beta) + [s.beta_lin]))
    if s.mu > 1e-6:
        if step_pred < 1 / math.sqrt(3):
            expon_used = 1.0
        else:
            expon_used = max(s.expon, 3.0 * step_pred ** 2)
    else:
        expon_used = max(1.0, min(s.expon, 3.0 * step_pred ** 2))
    tr = btrace(md.nlmi, s.Xn, s.Sn)
    if tr < 0:
        s.sigma = 0.8
    else:
        tmp1 = tr if md.nlmi > 0 else 0.0
        tmp2 = float(s.Xn_lin @ s.Sn_lin) if md.nlin > 0 else 0.0
        tmp12 = (tmp1 + tmp2) / (sum(md.msizes) + md.nlin)
        s.sigma = min(1.0, (tmp12 / s.mu) ** expon_used)
    return s.sigma


def corrector(s, halpha):
    """src/predictor_corrector.jl:181-246."""
    md = s.model
    s.predict = False
    t0 = time.perf_counter()
    h = s.Rp
    for i in range(md.nlmi):
        G = s.G[i]
        inner = G.T @ s.Rd[i] @ G + np.diag(s.D[i]) - np.diag((s.sigma * s.mu) / s.D[i]) - s.RNT[i]
        h = h + md.AA[i] @ my_kron(G, G, inner)
    if md.nlin > 0:
        tmp = (s.delX_lin * s.delS_lin) * s.Si_lin - (s.sigma * s.mu) * s.Si_lin
        h = h + md.C_lin @ ((s.X_lin * s.Si_lin) * s.Rd_lin + s.X_lin + tmp)
    _timed(s, "rhs", t0)
    if "rhs_corr" in s.hooks:
        s.hooks["rhs_corr"](s, h)
    if s.kit == 0:
        t0 = time.perf_counter()
        s.dely = _solve_direct(s, h)
        _timed(s, "schur_solve", t0)
    else:
        A = MyA(s)
        if s.preconditioner == 0:
            M = None
        elif s.preconditioner == 1:
            M = MyM(s, halpha)
        else:
            M = MyM_beta(halpha)
        t0 = time.perf_counter()
        s.dely, exit_code, num_iters = cg(A, h, tol=s.tol_cg, maxIter=10000, precon=M)
        _timed(s, "cg", t0)
        s.cg_iter_cor += num_iters
        s.cg_iter_tot += num_iters
    t0 = time.perf_counter()
    find_step(s)
    _timed(s, "find_step", t0)


def _steplen(mimi, tau):
    return 0.99 if mimi > -1e-6 else min(1.0, -tau / mimi)


def find_step(s):
    """src/predictor_corrector.jl:248-326."""
    md = s.model
    for i in range(md.nlmi):
        s.delS[i] = s.Rd[i] - mat(md.AA[i].T @ s.dely)
        Xi = my_kron(s.W[i], s.W[i], s.delS[i])
        if s.predict:
            s.delX[i] = mat(-vec(s.X[i]) - Xi)
        else:
            s.delX[i] = mat(vec((s.sigma * s.mu) * s.Si[i] - s.X[i]) - Xi + my_kron(s.G[i], s.G[i], s.RNT[i]))
        delSb = s.G[i].T @ s.delS[i] @ s.G[i]
        delXb = s.Gi[i] @ s.delX[i] @ s.Gi[i].T
        dd = s.DDsi[i]
        XXX = dd[None, :] * delXb * dd[:, None]
        XXX = (XXX + XXX.T) / 2
        s.alpha[i] = _steplen(eigmin(XXX), s.tau)
        XXX = dd[None, :] * delSb * dd[:, None]
        XXX = (XXX + XXX.T) / 2
        s.beta[i] = _steplen(eigmin(XXX), s.tau)
    if md.nlin > 0:
        find_step_lin(s)
    else:
        s.alpha_lin = 1.0
        s.beta_lin = 1.0
    if s.predict:
        for i in range(md.nlmi):
            s.Xn[i] = s.X[i] + s.alpha[i] * s.delX[i]
            s.Sn[i] = s.S[i] + s.beta[i] * s.delS[i]
            deed = s.D[i][:, None] + s.D[i][None, :]
            s.RNT[i] = -(s.Gi[i] @ s.delX[i] @ s.delS[i] @ s.G[i] + s.G[i].T @ s.delS[i] @ s.delX[i] @ s.Gi[i].T) / deed
    else:
        amin = min(list(s.alpha) + [s.alpha_lin])
        bmin = min(list(s.beta) + [s.beta_lin])
        s.yold = s.y
        s.y = s.y + bmin * s.dely
        for i in range(md.nlmi):
            s.X[i] = s.X[i] + amin * s.delX[i]
            s.X[i] = (s.X[i] + s.X[i].T) / 2
            s.S[i] = s.S[i] + bmin * s.delS[i]
            s.S[i] = (s.S[i] + s.S[i].T) / 2


def find_step_lin(s):
    """src/predictor_corrector.jl:329-364."""
    md = s.model
    s.delS_lin = s.Rd_lin - md.C_lin.T @ s.dely
    if s.predict:
        s.delX_lin = -s.X_lin - s.X_lin * s.Si_lin * s.delS_lin
    else:
        s.delX_lin = -s.X_lin - s.X_lin * s.Si_lin * s.delS_lin + (s.sigma * s.mu) * s.Si_lin + s.RNT_lin
    s.alpha_lin = _steplen(float((s.delX_lin / s.X_lin).min()), s.tau)
    s.beta_lin = _steplen(float((s.delS_lin / s.S_lin).min()), s.tau)
    if s.predict:
        s.Xn_lin = s.X_lin + s.alpha_lin * s.delX_lin
        s.Sn_lin = s.S_lin + s.beta_lin * s.delS_lin
        s.RNT_lin = -(s.delX_lin * s.delS_lin) * s.Si_lin
    else:
        amin = min(list(s.alpha) + [s.alpha_lin])
        bmin = min(list(s.beta) + [s.beta_lin])
        s.X_lin = s.X_lin + amin * s.delX_lin
        s.S_lin = s.S_lin + bmin * s.delS_lin
        s.S_lin_inv = 1.0 / s.S_lin


# --------------------------------------------------------------------------------------------
# driver  (src/Solvers.jl:304-361, :448-478)
# --------------------------------------------------------------------------------------------
def myIPstep(s, halpha):
    s.iter += 1
    if s.iter > s.maxit:
        s.status = 4
    s.cg_iter_pre = 0
    s.cg_iter_cor = 0
    find_mu(s)
    t0 = time.perf_counter()
    prepare_W(s)
    _timed(s, "prepare_W", t0)
    if "W" in s.hooks:
        s.hooks["W"](s)
    predictor(s, halpha)
    sigma_update(s)
    corrector(s, halpha)


def solve(s, halpha, max_iters=None):
    t1 = time.perf_counter()
    setup_solver(s, halpha)
    initial_point(s)
    while s.status == 0:
        t2 = time.perf_counter()
        myIPstep(s, halpha)
        s.itertime = time.perf_counter() - t2
        s.tol_cg = max(s.tol_cg * s.tol_cg_up, s.tol_cg_min)
        check_convergence(s)
        s.trace.append(dict(iter=s.iter, obj=s.primal_obj, dual_obj=s.dual_obj, dimacs=s.DIMACS_error, mu=s.mu,
                            sigma=s.sigma, alpha=min(list(s.alpha) + [s.alpha_lin]),
                            beta=min(list(s.beta) + [s.beta_lin]), err=[s.err1, s.err2, s.err3, s.err4, s.err5, s.err6],
                            cg_pre=s.cg_iter_pre, cg_cor=s.cg_iter_cor, itertime=s.itertime))
        if s.verb > 0:
            print(f"{s.iter:3d} {s.primal_obj:16.8e} {s.DIMACS_error:9.2e} {s.cg_iter_pre + s.cg_iter_cor:7d} {s.itertime:8.3f}")
        if s.preconditioner == 4:
            if (s.cg_iter_cor / 2 > s.erank * s.model.nlmi * math.sqrt(s.model.n) / 20 and s.iter > math.sqrt(s.model.n) / 60) \
                    or s.cg_iter_cor > 100:
                s.preconditioner = 1
                s.aamat = 2
        if max_iters is not None and s.iter >= max_iters:
            break
    s.tottime = time.perf_counter() - t1
    return s


def solve_raw(raw, options):
    o = dict(DEFAULT_OPTIONS)
    o.update(options)
    md = prepare_model(raw, datarank=int(o["datarank"]), kappa=int(o["datasparsity"]))
    s, ha = load(md, o)
    return solve(s, ha, max_iters=options.get("_max_iters"))
