"""The reference's JuMP examples (examples/*.jl) written out in the form `MOI.copy_to` hands them to Loraine after the
MathOptInterface bridges: free variables, one VAF-in-PSD constraint per PSD variable/constraint, every scalar
(in)equality as rows of one VAF-in-Nonnegatives block (an equality is two opposite inequalities).  The optimum does not
depend on the row order chosen by the bridges.

Each builder returns a dict of RawProblem fields (n, msizes, A, b, b_const, C_lin, d_lin) + `max_sense`."""
import numpy as np
import scipy.sparse as sp


def _psd_variable_block(m, var_of):
    """PSD variable X (m x m): A[.,k+1] = E_k for the variable holding X[i,j]; A[.,1] = 0."""
    k, p, q, v = [], [], [], []
    for j in range(m):
        for i in range(j + 1):
            kk = var_of(i, j) + 1
            k.append(kk); p.append(i); q.append(j); v.append(1.0)
            if i != j:
                k.append(kk); p.append(j); q.append(i); v.append(1.0)
    return dict(k=np.array(k), p=np.array(p), q=np.array(q), v=np.array(v, dtype=float))


def _lin(rows, n):
    """rows: list of (coeff dict var->value, constant) meaning coeff.x + constant >= 0.
    C_lin = -coeff' (src/MOI_wrapper.jl:149), d_lin = constants (:217)."""
    R, Cc, V, d = [], [], [], []
    for r, (co, const) in enumerate(rows):
        for var, val in co.items():
            R.append(r); Cc.append(var); V.append(val)
        d.append(const)
    coeff = sp.csr_matrix((V, (R, Cc)), shape=(len(rows), n))
    return (-coeff.T).tocsc(), np.array(d, dtype=float)


def _tri(i, j):
    i, j = min(i, j), max(i, j)
    return j * (j + 1) // 2 + i


def ex_corr(sense):
    """examples/ex_corr.jl:9-31; sense = 'Max' -> 0.8719210472, 'Min' -> -0.9779977649 (value of rho_AC)."""
    n = 6
    rows = []
    for i in range(3):
        rows += [({_tri(i, i): 1.0}, -1.0), ({_tri(i, i): -1.0}, 1.0)]           # rho_ii == 1
    rows += [({_tri(0, 1): 1.0}, 0.2), ({_tri(0, 1): -1.0}, -0.1)]             # -0.2 <= rho_AB <= -0.1
    rows += [({_tri(1, 2): 1.0}, -0.4), ({_tri(1, 2): -1.0}, 0.5)]             # 0.4 <= rho_BC <= 0.5
    C_lin, d_lin = _lin(rows, n)
    b0 = np.zeros(n); b0[_tri(0, 2)] = 1.0
    max_sense = sense == "Max"
    return dict(n=n, msizes=[3], A=[_psd_variable_block(3, _tri)], b=b0 if max_sense else -b0, b_const=0.0,
                C_lin=C_lin, d_lin=d_lin, max_sense=max_sense)


def ex_dist():
    """examples/ex_dist.jl:8-40: min c2 s.t. D_ij^2 <= Q_ii+Q_jj-2Q_ij <= c2 D_ij^2, Q PSD 4x4, Q_11 == 0, c2 >= 1.  Optimum 4/3."""
    D = np.array([[0, 1, 1, 1], [1, 0, 2, 2], [1, 2, 0, 2], [1, 2, 2, 0]], dtype=float)
    n = 11                                   # variable 0 = c2, variables 1..10 = triangle of Q
    q = lambda i, j: 1 + _tri(i, j)
    rows = [({0: 1.0}, -1.0)]                # c2 >= 1
    for i in range(4):
        for j in range(i + 1, 4):
            e = {q(i, i): 1.0, q(j, j): 1.0, q(i, j): -2.0}
            rows.append((dict(e), -D[i, j] ** 2))
            e2 = {k: -v for k, v in e.items()}
            e2[0] = D[i, j] ** 2
            rows.append((e2, 0.0))
    rows += [({q(0, 0): 1.0}, 0.0), ({q(0, 0): -1.0}, 0.0)]                   # fix(Q[1,1], 0)
    C_lin, d_lin = _lin(rows, n)
    b0 = np.zeros(n); b0[0] = 1.0
    return dict(n=n, msizes=[4], A=[_psd_variable_block(4, lambda i, j: q(i, j))], b=-b0, b_const=0.0, C_lin=C_lin,
                d_lin=d_lin, max_sense=False)


def ex_maxcut4():
    """examples/ex_maxcut.jl:18-47: max 0.25 <L, X>, diag(X) == 1, X PSD 4x4.  Optimum 17, X = xx', x = (1,-1,-1,1)."""
    W = np.array([[0, 1, 5, 0], [1, 0, 0, 9], [5, 0, 0, 2], [0, 9, 2, 0]], dtype=float)
    L = np.diag(W.sum(axis=1)) - W
    n = 10
    rows = []
    for i in range(4):
        rows += [({_tri(i, i): 1.0}, -1.0), ({_tri(i, i): -1.0}, 1.0)]
    C_lin, d_lin = _lin(rows, n)
    b0 = np.zeros(n)
    for j in range(4):
        for i in range(j + 1):
            b0[_tri(i, j)] = 0.25 * L[i, j] * (1.0 if i == j else 2.0)
    return dict(n=n, msizes=[4], A=[_psd_variable_block(4, _tri)], b=b0, b_const=0.0, C_lin=C_lin, d_lin=d_lin,
                max_sense=True)


def ex_k_lp():
    """examples/k.jl:17-38 (Float64 version): max 2x, 1 <= x <= 2 -> objective 4, x = 2."""
    C_lin, d_lin = _lin([({0: 1.0}, -1.0), ({0: -1.0}, 2.0)], 1)
    return dict(n=1, msizes=[], A=[], b=np.array([2.0]), b_const=0.0, C_lin=C_lin, d_lin=d_lin, max_sense=True)


def objective_value(spec, y):
    val = float(np.dot(spec["b"], y)) - spec["b_const"]        # src/MOI_wrapper.jl:315-319
    return val if spec["max_sense"] else -val


def fields(spec):
    return {k: spec[k] for k in ("n", "msizes", "A", "b", "b_const", "C_lin", "d_lin")}
