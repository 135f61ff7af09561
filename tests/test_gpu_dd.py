"""GPU parity tests of the double-double LP path (include/loraine_b200_dd.h) against the extended-precision
oracle (oracle/dd_lp_oracle.py, 160-bit mpmath = "exact" next to the ~106 bits of Float64x2).  Tolerances: a double-double
operation is good to ~2^-104 = 5e-32; sums of length n and the conditioning of the Schur matrix leave 1e-26 relative."""
from fractions import Fraction

import mpmath as mp
import numpy as np
import pytest

import jump_examples as je
from dd_common import random_lp
from oracle import dd_lp_oracle as ddo
from oracle import loraine_oracle as lo
from oracle import sdpa_io

pytestmark = pytest.mark.gpu
TOL = mp.mpf(1e-26)


def _mp_vec(hi, lo):
    return [mp.mpf(float(a)) + mp.mpf(float(b)) for a, b in zip(np.ravel(hi, order="F"), np.ravel(lo, order="F"))]


def _rel(dev, ref):
    num = mp.sqrt(sum((a - mp.mpf(b)) ** 2 for a, b in zip(dev, ref)))
    den = mp.sqrt(sum(mp.mpf(b) ** 2 for b in ref))
    return num / den if den != 0 else num


def _pair(p):
    return mp.mpf(float(p[0])) + mp.mpf(float(p[1]))


def _setup(pkg, spec, eD=1e-25):
    from loraine_jl_b200 import dd_lp
    raw = pkg.RawProblem(**{k: spec[k] for k in ("n", "msizes", "A", "b", "b_const", "C_lin", "d_lin")})
    md = pkg.prepare_model(raw)
    o = dict(pkg.DEFAULT_OPTIONS, eDIMACS=eD, verb=0)
    s = dd_lp.DDSolver(md, o)
    dd_lp.setup_solver(s)
    dd_lp.initial_point(s)
    oraw = sdpa_io.RawProblem(**{k: spec[k] for k in ("n", "msizes", "A", "b", "b_const", "C_lin", "d_lin")})
    omd = lo.prepare_model(oraw, 0, 8)
    return dd_lp, s, omd


@pytest.mark.parametrize("n,nlin", [(12, 30), (70, 160)])
def test_phase_parity_two_iterations(pkg, n, nlin):
    spec = random_lp(n, nlin, 11 + n)
    dd_lp, s, omd = _setup(pkg, spec)
    o = ddo.setup(omd.C_lin, omd.d_lin, omd.b, dict(eDIMACS=1e-25), prec=160, promote_all=True)
    ddo.initial_point(o)
    for it in range(2):
        dd_lp.find_mu(s); ddo.find_mu(o)
        assert abs(_pair(s.mu) - o.mu) <= TOL * abs(o.mu)
        dd_lp.prepare_W(s); ddo.prepare_W(o)
        assert _rel(_mp_vec(*s.get_array("SI")), o.Si) <= TOL
        # predictor, call by call (the library keeps H until the corrector has updated the iterate)
        s.predict = True
        s._call("lrn_dd_residuals")
        s._call("lrn_dd_schur_assemble")
        s._call("lrn_dd_rhs_predictor")
        assert s._call("lrn_dd_schur_factor", allow_positive=True) == 0
        ddo.predictor(o)
        assert _rel(_mp_vec(*s.get_array("RP")), o.Rp) <= TOL
        assert _rel(_mp_vec(*s.get_array("RD")), o.Rd) <= TOL
        Hh, Hl = s.get_array("H")
        assert _rel(_mp_vec(Hh, Hl), [o.H[i][j] for j in range(n) for i in range(n)]) <= TOL
        assert _rel(_mp_vec(*s.get_array("RHS")), o.h_pred) <= TOL
        Lh, Ll = s.get_array("L")
        Lh, Ll = np.tril(Lh), np.tril(Ll)
        assert _rel(_mp_vec(Lh, Ll), [o.L[i][j] for j in range(n) for i in range(n)]) <= TOL
        s.chol_is_factor_object = False
        s._call("lrn_dd_schur_solve", 3)
        assert _rel(_mp_vec(*s.get_array("DELY")), o.dely_pred) <= TOL
        dd_lp._find_step(s, True)
        # (the oracle's predictor has already run find_step_lin: its delX / delS / Xn / Sn / RNT are the predictor's)
        for name, ref in (("DELX", o.delX), ("DELS", o.delS), ("XN", o.Xn), ("SN", o.Sn), ("RNT", o.RNT)):
            assert _rel(_mp_vec(*s.get_array(name)), ref) <= TOL, name
        assert abs(_pair(s.alpha_lin) - o.alpha) <= TOL and abs(_pair(s.beta_lin) - o.beta) <= TOL
        dd_lp.sigma_update(s); ddo.sigma_update(o)
        assert abs(s.sigma - float(o.sigma)) <= 1e-14 * float(o.sigma)
        o.sigma = mp.mpf(s.sigma)                          # same Float64 parameter on both sides from here on
        dd_lp.corrector(s); ddo.corrector(o)
        assert _rel(_mp_vec(*s.get_array("RHS")), o.h_corr) <= TOL
        assert _rel(_mp_vec(*s.get_array("DELY")), o.dely) <= TOL
        dd_lp.get_solution(s)
        assert _rel(_mp_vec(*s.y_dd), o.y) <= TOL
        assert _rel(_mp_vec(*s.X_lin_dd), o.X) <= TOL
        assert _rel(_mp_vec(*s.S_lin_dd), o.S) <= TOL
        s.itertime = 0.0
        dd_lp.check_convergence(s); ddo.check_convergence(o)
        for dev, ref in ((s.err2, o.err2), (s.err3, o.err3), (s.err4, o.err4), (s.err5, o.err5), (s.err6, o.err6)):
            d = mp.mpf(dev.numerator) / mp.mpf(dev.denominator)
            # the denominators 1 + norm(b), 1 + norm(d_lin) are Float64 numbers in the reference (Float64 model data): the
            # oracle's and the library's differ by an ulp of Float64 at most
            assert abs(d - ref) <= mp.mpf(1e-15) * (abs(ref) + mp.mpf(1e-30))
    s.close()


def test_k_lp_float64x2_end_to_end(pkg):
    """examples/k.jl:8-38: Model(Loraine.Optimizer{Float64x2}); max 2x, 1 <= x <= 2 -> objective 4, x = 2 (here to 1e-24)."""
    spec = je.ex_k_lp()
    opt = pkg.Optimizer(T="Float64x2")
    opt.set_attribute("verb", 0)
    opt.set_attribute("eDIMACS", 1e-25)
    opt.copy_to(pkg.RawProblem(**je.fields(spec)), max_sense=True)
    opt.optimize()
    s = opt.solver
    assert opt.termination_status() == "OPTIMAL"
    assert abs(opt.objective_value_dd() - 4) <= Fraction(1, 10 ** 24)
    assert abs(opt.objective_value() - 4) <= 1e-15
    x = Fraction(float(s.y_dd[0][0])) + Fraction(float(s.y_dd[1][0]))
    assert abs(x - 2) <= Fraction(1, 10 ** 24)
    assert abs(s.X_lin[0]) <= 1e-24 and abs(s.X_lin[1] - 2) <= 1e-15           # shadow prices 0 and 2
    omd = lo.prepare_model(sdpa_io.RawProblem(**je.fields(spec)), 0, 8)
    ref = ddo.solve(omd.C_lin, omd.d_lin, omd.b, dict(eDIMACS=1e-25), prec=106)   # the reference's mixed first iteration
    assert ref.status == 1 and abs(s.iter - ref.iter) <= 1
    assert s.DIMACS_error < Fraction(1, 10 ** 25)


def test_random_lp_end_to_end(pkg):
    spec = random_lp(40, 100, 5)
    dd_lp, s, omd = _setup(pkg, spec, eD=1e-24)
    dd_lp.solve(s, setup=False)
    ref = ddo.solve(omd.C_lin, omd.d_lin, omd.b, dict(eDIMACS=1e-24), prec=160, promote_all=True)
    assert s.status == 1 and ref.status == 1 and abs(s.iter - ref.iter) <= 1
    if s.iter == ref.iter:
        assert _rel(_mp_vec(*s.y_dd), ref.y) <= mp.mpf(1e-20)
    by = _pair(s.by)
    assert abs(by - ref.by) <= mp.mpf(1e-22) * abs(ref.by)
    assert s.DIMACS_error < Fraction(1, 10 ** 24)
    # a Float64 solve of the same LP stalls around 1e-16: the extra digits are real
    opt = pkg.Optimizer()
    opt.set_attribute("verb", 0)
    opt.copy_to(pkg.RawProblem(**{k: spec[k] for k in ("n", "msizes", "A", "b", "b_const", "C_lin", "d_lin")}))
    opt.optimize()
    assert abs(float(opt.solver.model.b @ opt.solver.y) - float(ref.by)) <= 1e-6 * abs(float(ref.by))
    s.close()


def test_regularised_retry_and_many_tiles(pkg):
    """n = 300 (10 Cholesky tiles, ragged last one): L L' = H and H dely = h at Float64 resolution (indexing), then the
    shift + refactor path (src/predictor_corrector.jl:66-88) with the H^-1 H^-1 h quirk of the `Cholesky` object."""
    n, nlin = 300, 700
    spec = random_lp(n, nlin, 9, density=0.1)
    dd_lp, s, omd = _setup(pkg, spec)
    dd_lp.find_mu(s); dd_lp.prepare_W(s)
    s._call("lrn_dd_residuals"); s._call("lrn_dd_schur_assemble"); s._call("lrn_dd_rhs_predictor")
    assert s._call("lrn_dd_schur_factor", allow_positive=True) == 0
    H = s.get_array("H")[0]
    L = np.tril(s.get_array("L")[0])
    M = omd.C_lin.toarray()
    assert np.linalg.norm(H - M @ M.T) <= 1e-13 * np.linalg.norm(H)          # x = s = 1 at the initial point
    assert np.linalg.norm(L @ L.T - H) <= 1e-13 * np.linalg.norm(H)
    s._call("lrn_dd_schur_solve", 3)
    h = s.get_array("RHS")[0]
    y3 = s.get_array("DELY")[0]
    assert np.linalg.norm(H @ y3 - h) <= 1e-11 * np.linalg.norm(h)
    s._call("lrn_dd_schur_solve", 6)
    y6 = s.get_array("DELY")[0]
    assert np.linalg.norm(H @ (H @ y6) - h) <= 1e-9 * np.linalg.norm(h)
    # an indefinite matrix: shift the diagonal far down, the factorisation must report a pivot, shifting back must repair it
    s._call("lrn_dd_schur_shift", -1e6)
    assert s._call("lrn_dd_schur_factor", allow_positive=True) > 0
    s._call("lrn_dd_schur_shift", 1e6)
    assert s._call("lrn_dd_schur_factor", allow_positive=True) == 0
    t = s.timers()
    print("dd n=300: assemble %.2f ms, factor %.2f ms, solve %.2f ms" % (t["schur_assemble"], t["schur_factor"], t["schur_solve"]))
    s.close()
