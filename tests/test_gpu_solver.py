"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on the same inputs, function by function
and end to end, plus the committed golden fixtures and the reference's known answers."""
import os

import numpy as np
import pytest

import jump_examples as je

pytestmark = pytest.mark.gpu

OPTS_SDPA = dict(kit=0, tol_cg=1e-2, tol_cg_min=1e-6, eDIMACS=1e-6, preconditioner=1, erank=1, aamat=2, verb=0, datarank=0,
                 initpoint=1, maxit=100, datasparsity=8)


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return z, (int(z["n"]), [int(b) for b in z["bs"]], z["c"], z["body"])


def make_pair(pkg, arrays, opts):
    """the same problem prepared for the CUDA path and for the oracle"""
    from oracle import loraine_oracle as lo, sdpa_io
    opt = pkg.Optimizer()
    for k, v in opts.items():
        opt.set_attribute(k, v)
    if isinstance(arrays, dict):
        opt.copy_to(pkg.RawProblem(**je.fields(arrays)), max_sense=arrays["max_sense"])
        oraw = sdpa_io.RawProblem(**je.fields(arrays))
    else:
        opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
        oraw = sdpa_io.raw_from_sdpa_arrays(*arrays)
    o = dict(lo.DEFAULT_OPTIONS)
    o.update(opts)
    md = lo.prepare_model(oraw, datarank=int(o["datarank"]), kappa=int(o["datasparsity"]))
    s, ha = lo.load(md, o)
    return opt, (lo, s, ha)


def step_both(pkg, opt, ora, iters):
    """run `iters` full IP iterations on both sides"""
    from loraine_jl_b200 import solver as S
    lo, s, ha = ora
    g = opt.solver
    S.setup_solver(g, opt.halpha)
    S.initial_point(g)
    lo.setup_solver(s, ha)
    lo.initial_point(s)
    for _ in range(iters):
        S.myIPstep(g, opt.halpha)
        g.itertime = 0.0
        g.tol_cg = max(g.tol_cg * g.tol_cg_up, g.tol_cg_min)
        S.check_convergence(g)
        lo.myIPstep(s, ha)
        s.tol_cg = max(s.tol_cg * s.tol_cg_up, s.tol_cg_min)
        lo.check_convergence(s)
    return g, s


@pytest.mark.parametrize("name,nblk", [("theta1", 1), ("control1", 2), ("vib3", 2)])
def test_first_iterations_match_oracle(pkg, golden_dir, name, nblk):
    z, arrays = golden(golden_dir, name)
    opt, ora = make_pair(pkg, arrays, OPTS_SDPA)
    g, s = step_both(pkg, opt, ora, 3)
    # the fourth iteration is compared phase by phase
    from loraine_jl_b200 import solver as S
    lo = ora[0]
    for side, mod, st, ha in ((0, S, g, opt.halpha), (1, lo, s, ora[2])):
        st.iter += 1
        st.cg_iter_pre = st.cg_iter_cor = 0
        mod.find_mu(st)
        mod.prepare_W(st)
    assert abs(g.mu - s.mu) <= 1e-10 * abs(s.mu)
    for i in range(nblk):
        assert relerr(g.get_array("W", i), s.W[i]) <= 1e-9                 # W is unique (W S W = X)
        assert relerr(np.sort(g.get_array("D", i)), np.sort(s.D[i])) <= 1e-9
        Gg, Gig = g.get_array("G", i), g.get_array("GI", i)
        m = Gg.shape[0]
        assert np.linalg.norm(Gg @ Gig - np.eye(m)) <= 1e-7 * m            # Gi = inv(G) in closed form
        assert relerr(g.get_array("SI", i), s.Si[i]) <= 1e-8
    S.predictor(g, opt.halpha)
    lo.predictor(s, ora[2])
    H = g.get_array("H")
    got = []
    s.hooks["H"] = lambda s_, Hm: got.append(Hm.copy())
    # oracle H of the same iterate (recompute from the oracle's W)
    Ho = lo.makeBBBBs(s.model, s.W)
    if s.model.nlin > 0:
        Ho = Ho + lo.lp_schur(s.model, s.X_lin * s.S_lin_inv)
    assert relerr(H, Ho) <= 1e-9                                           # oracle W differs by its own SVD rounding
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-7
    assert np.allclose(g.alpha, s.alpha, rtol=1e-6) and np.allclose(g.beta, s.beta, rtol=1e-6)
    assert abs(S.sigma_update(g) - lo.sigma_update(s)) <= 1e-6
    S.corrector(g, opt.halpha)
    lo.corrector(s, ora[2])
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-6
    S.check_convergence(g)
    lo.check_convergence(s)
    assert abs(g.DIMACS_error - s.DIMACS_error) <= 1e-6 * max(1.0, s.DIMACS_error)
    g.close()


@pytest.mark.parametrize("datarank", [-1, 0])
def test_sparse_data_forms_match_oracle(pkg, datarank):
    """m = 192 max-cut instance: the stored positions are < 5 % of m^2, so the RHS congruences are sampled at the stored
    positions, G'RdG is shared across the iteration and G'(MG) is a gather product (lrn_rhs_*, lrn_find_step); the iterates
    must follow the oracle, which evaluates the reference's dense formulas (src/predictor_corrector.jl:183-192, :248-326)."""
    from loraine_jl_b200 import solver as S
    arrays = pkg.problems.maxcut_torus(12, 16, 11)
    opt, ora = make_pair(pkg, arrays, dict(kit=0, datarank=datarank, initpoint=1, verb=0))
    g, s = step_both(pkg, opt, ora, 3)
    lo = ora[0]
    y, X, _ = S.get_solution(g)
    assert relerr(y, s.y) <= 1e-7
    assert relerr(X[0], s.X[0]) <= 1e-7
    for side, mod, st, ha in ((0, S, g, opt.halpha), (1, lo, s, ora[2])):
        st.iter += 1
        mod.find_mu(st)
        mod.prepare_W(st)
        mod.predictor(st, ha)
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-7
    assert np.allclose(g.alpha, s.alpha, rtol=1e-6) and np.allclose(g.beta, s.beta, rtol=1e-6)
    assert abs(S.sigma_update(g) - lo.sigma_update(s)) <= 1e-6
    S.corrector(g, opt.halpha)
    lo.corrector(s, ora[2])
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-6
    g.close()


def test_schur_matrix_parity_1e11(pkg, golden_dir):
    """north_star: assembled Schur matrix within 1e-11 relative Frobenius error for the SAME W (general + LP + rank-one)."""
    from oracle import loraine_oracle as lo
    from loraine_jl_b200 import solver as S
    for name, opts in (("theta1", OPTS_SDPA), ("control1", OPTS_SDPA), ("tru3", OPTS_SDPA),
                       ("control1", dict(OPTS_SDPA, schur_split=1, datasparsity=8))):
        z, arrays = golden(golden_dir, name)
        opt, ora = make_pair(pkg, arrays, opts)
        g, s = step_both(pkg, opt, ora, 2)
        S.find_mu(g); S.prepare_W(g)
        g._call("lrn_residuals"); g._call("lrn_schur_assemble")
        H = g.get_array("H")
        Wg = [g.get_array("W", i) for i in range(s.model.nlmi)]
        Ho = lo.makeBBBBs(s.model, Wg)
        if s.model.nlin > 0:
            y, X, xl = S.get_solution(g)
            sl = np.zeros(s.model.nlin)
            import ctypes as C
            g._call("lrn_get_slack", None, sl.ctypes.data_as(C.POINTER(C.c_double)))
            Ho = Ho + lo.lp_schur(s.model, xl / sl)
        assert relerr(H, Ho) <= 1e-11, name
        g.close()
    # rank-one path against the oracle's makeBBBB_rank1 with the device's G
    arrays = pkg.problems.maxcut_torus(8, 12, 96)
    o = dict(OPTS_SDPA, datarank=-1)
    opt, ora = make_pair(pkg, arrays, o)
    g, s = step_both(pkg, opt, ora, 2)
    S.find_mu(g); S.prepare_W(g)
    g._call("lrn_residuals"); g._call("lrn_schur_assemble")
    H = g.get_array("H")
    Ho = lo.makeBBBB_rank1(s.model.n, 1, s.model.B, [g.get_array("G", 0)])
    assert relerr(H, Ho) <= 1e-11
    # ... and the rank-one path equals the general path
    opt2, ora2 = make_pair(pkg, arrays, dict(OPTS_SDPA, datarank=0))
    g2, s2 = step_both(pkg, opt2, ora2, 2)
    S.find_mu(g2); S.prepare_W(g2)
    g2._call("lrn_residuals"); g2._call("lrn_schur_assemble")
    assert relerr(g2.get_array("H"), H) <= 1e-10
    g.close(); g2.close()


@pytest.mark.parametrize("name", ["theta1", "control1", "tru3", "vib3"])
def test_end_to_end_sdplib(pkg, golden_dir, name):
    """objective within eDIMACS of the oracle / golden fixture, IP iteration count within +-1"""
    z, arrays = golden(golden_dir, name)
    opt = pkg.Optimizer()
    for k, v in OPTS_SDPA.items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    opt.optimize()
    s = opt.solver
    assert s.status == 1
    assert abs(s.iter - int(z["oracle_iters"])) <= 1
    assert abs(s.primal_obj - float(z["oracle_obj"])) <= 1e-6 * (1 + abs(float(z["oracle_obj"])))
    assert abs(s.dual_obj - float(z["oracle_dual_obj"])) <= 1e-5 * (1 + abs(float(z["oracle_obj"])))
    if name == "theta1":
        assert abs(opt.objective_value() - 23) <= 23e-6                    # examples/solve_sdpa.jl:61
    s.close()


def test_reference_jump_examples(pkg):
    """the reference's own end-to-end assertions (examples/*.jl) through the CUDA path"""
    for sense, want in (("Max", 0.8719210472), ("Min", -0.9779977649)):
        spec = je.ex_corr(sense)
        opt = pkg.Optimizer(); opt.set_attribute("verb", 0)
        opt.copy_to(pkg.RawProblem(**je.fields(spec)), max_sense=spec["max_sense"])
        opt.optimize()
        assert opt.termination_status() == "OPTIMAL"
        assert abs(opt.objective_value() - want) <= 1e-6 * abs(want)
        opt.solver.close()
    spec = je.ex_dist()
    opt = pkg.Optimizer(); opt.set_attribute("verb", 0)
    opt.copy_to(pkg.RawProblem(**je.fields(spec)), max_sense=False)
    opt.optimize()
    assert opt.termination_status() == "OPTIMAL" and abs(opt.objective_value() - 4 / 3) <= 1e-4
    opt.solver.close()
    spec = je.ex_maxcut4()
    opt = pkg.Optimizer(); opt.set_attribute("verb", 0)
    opt.copy_to(pkg.RawProblem(**je.fields(spec)), max_sense=True)
    opt.optimize()
    assert abs(opt.objective_value() - 17) <= 17e-5
    opt.solver.close()
    spec = je.ex_k_lp()                                                     # pure LP block, nlmi = 0
    opt = pkg.Optimizer(); opt.set_attribute("verb", 0)
    opt.copy_to(pkg.RawProblem(**je.fields(spec)), max_sense=True)
    opt.optimize()
    assert abs(opt.objective_value() - 4) <= 4e-6 and abs(opt.solver.y[0] - 2) <= 2e-6
    opt.solver.close()


@pytest.mark.parametrize("cfg", ["C2-mini", "C4-mini", "C5-mini"])
def test_synthetic_minis_direct(pkg, cfg):
    from oracle import loraine_oracle as lo, sdpa_io
    c = pkg.problems.CONFIGS[cfg]
    arrays = c["gen"]()
    o = dict(c["options"], verb=0)
    ref = lo.solve_raw(sdpa_io.raw_from_sdpa_arrays(*arrays), o)
    opt = pkg.Optimizer()
    for k, v in o.items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    opt.optimize()
    s = opt.solver
    assert s.status == ref.status == 1
    assert abs(s.iter - ref.iter) <= 1
    assert abs(s.primal_obj - ref.primal_obj) <= 1e-6 * (1 + abs(ref.primal_obj))
    np.testing.assert_allclose(s.y, ref.y, atol=1e-4 * (1 + np.abs(ref.y).max()))
    s.close()


@pytest.mark.parametrize("prec", [0, 1, 2, 4])
def test_cg_path_theta(pkg, golden_dir, prec):
    """kit = 1: objective and CG / IP iteration counts against the oracle (parity unpinned by the reference)."""
    from oracle import loraine_oracle as lo, sdpa_io
    arrays = pkg.problems.theta_torus(6, 8)
    o = dict(pkg.problems.CONFIGS["C3-mini"]["options"], verb=0, preconditioner=prec)
    ref = lo.solve_raw(sdpa_io.raw_from_sdpa_arrays(*arrays), o)
    opt = pkg.Optimizer()
    for k, v in o.items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    opt.optimize()
    s = opt.solver
    assert s.status == 1
    assert abs(s.primal_obj - 24.0) <= 1e-3                                # theta of the 6 x 8 torus is 24
    assert abs(s.primal_obj - ref.primal_obj) <= 1e-4 * (1 + abs(ref.primal_obj))
    assert abs(s.iter - ref.iter) <= 1
    assert abs(s.cg_iter_tot - ref.cg_iter_tot) <= max(10, 0.15 * ref.cg_iter_tot)
    s.close()


@pytest.mark.parametrize("gen,args", [("theta_torus", (6, 8)), ("multiblock_lp", (3, 12, 10, 7)), ("maxcut_torus", (6, 8, 3))])
def test_cg_operator_sparse_and_dense_forms(pkg, gen, args):
    """MyA (src/Solvers.jl:572-614): the sparse-aware operator (Z = M W by gathers, <calA_j, W Z> sampled) and the dense
    W M W form give the oracle's result on the same iterate."""
    from oracle import loraine_oracle as lo
    from loraine_jl_b200 import solver as S
    import ctypes as C
    arrays = getattr(pkg.problems, gen)(*args)
    o = dict(kit=1, preconditioner=0, initpoint=1, verb=0)
    opt, ora = make_pair(pkg, arrays, o)
    g, s = step_both(pkg, opt, ora, 2)
    for mod, st in ((S, g), (lo, s)):
        st.iter += 1; st.cg_iter_pre = st.cg_iter_cor = 0
        mod.find_mu(st); mod.prepare_W(st)
    x = np.random.default_rng(7).standard_normal(s.model.n)
    ref = lo.MyA(s)(x)
    dpx = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    outs = []
    for mode in (1.0, 0.0):
        g._call("lrn_set_option", b"sparse_op", mode)
        out = np.zeros_like(x)
        g._call("lrn_apply_operator", -1, dpx(x), dpx(out))
        assert relerr(out, ref) <= 1e-8
        outs.append(out)
    assert relerr(outs[0], outs[1]) <= 1e-11
    g.close()


def test_cg_operator_and_preconditioner_apply(pkg):
    """MyA and MyM applied to a random vector vs the oracle's functors on the same iterate (LP block included)."""
    from oracle import loraine_oracle as lo
    from loraine_jl_b200 import solver as S
    import ctypes as C
    arrays = pkg.problems.multiblock_lp(3, 12, 10, 7)
    o = dict(kit=1, preconditioner=1, erank=1, aamat=2, initpoint=1, verb=0, eDIMACS=1e-6)
    opt, ora = make_pair(pkg, arrays, o)
    g, s = step_both(pkg, opt, ora, 2)
    for mod, st in ((S, g), (lo, s)):
        st.iter += 1; st.cg_iter_pre = st.cg_iter_cor = 0
        mod.find_mu(st); mod.prepare_W(st)
    g._call("lrn_residuals"); g._call("lrn_rhs_predictor"); g._call("lrn_prec_prepare", 1)
    ha = ora[2]
    lo.Prec_for_CG_tilS_prep(s, ha)
    x = np.random.default_rng(1).standard_normal(s.model.n)
    out = np.zeros_like(x)
    dpx = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    g._call("lrn_apply_operator", -1, dpx(x), dpx(out))
    assert relerr(out, lo.MyA(s)(x)) <= 1e-8
    g._call("lrn_apply_operator", 1, dpx(x), dpx(out))
    assert relerr(out, lo.MyM(s, ha)(x)) <= 1e-6
    g.close()


def test_regularisation_path_matches_oracle(pkg):
    """Linearly dependent constraint matrices make H singular: the reference catches PosDefException, adds 1e-4 I until
    `isposdef` (src/predictor_corrector.jl:59-85) and then solves with the `Cholesky` object (H^-1 H^-1 h quirk, :89-90).
    The CUDA path must take the same branch (lrn_schur_factor > 0 -> lrn_schur_shift loop -> lrn_schur_solve(6))."""
    from oracle import loraine_oracle as lo, sdpa_io
    from loraine_jl_b200 import solver as S
    n, bs, c, body = pkg.problems.large_schur(12, 20, 3)
    body = np.array(body)
    dup = body[body[:, 0] == 2].copy()
    body = body[body[:, 0] != 3]
    dup[:, 0] = 3                                   # F_3 := F_2  -> H has two identical rows/columns
    body = np.concatenate([body, dup])
    c = np.array(c); c[2] = c[1]
    opts = dict(kit=0, initpoint=1, eDIMACS=1e-6, verb=0, maxit=3)
    opt, ora = make_pair(pkg, (n, bs, c, body), opts)
    g, s = step_both(pkg, opt, ora, 1)
    assert s.regcount == 1 and g.regcount == 1
    assert s.chol_is_factor_object and g.cholBBBB.is_cholesky_object
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-6
    g.close()


def test_non_pd_iterate_is_regularised_like_try_cholesky(pkg):
    """src/prepare_W.jl:5-26: X not positive definite -> X += 1e-5 I until the factorisation succeeds."""
    from loraine_jl_b200 import solver as S
    arrays = pkg.problems.theta_torus(3, 4)
    opt = pkg.Optimizer()
    for k, v in dict(kit=0, initpoint=1, verb=0).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    g = opt.solver
    S.setup_solver(g, opt.halpha)
    S.initial_point(g)
    m, n = g.model.msizes[0], g.model.n
    X = np.eye(m); X[0, 0] = -2e-5                  # needs 3 shifts of 1e-5
    S.set_iterate(g, [X], [2.0 * np.eye(m)], np.zeros(n), np.zeros(0), np.zeros(0))
    S.find_mu(g)
    S.prepare_W(g)
    assert g.status == 0
    y, Xd, _ = S.get_solution(g)
    assert abs(Xd[0][0, 0] - (-2e-5 + 3e-5)) <= 1e-12 and abs(Xd[0][1, 1] - (1 + 3e-5)) <= 1e-12
    W = g.get_array("W", 0)
    assert np.all(np.isfinite(W)) and np.linalg.eigvalsh(W)[0] > 0
    g.close()


def test_cg_path_with_lp_block_iterations(pkg):
    """kit = 1 with an LP block: AAAATtau is not diagonal (dense + Cholesky on the device, src/Solvers.jl:743-745,874,900).
    Three full IP iterations (predictor + corrector PCG solves) track the oracle; in the fourth both PCGs break down
    (exit code -13) on this instance, after which the iterates are rounding-dependent."""
    arrays = pkg.problems.multiblock_lp(3, 12, 10, 7)
    o = dict(kit=1, preconditioner=1, erank=1, aamat=2, initpoint=1, verb=0, eDIMACS=1e-12, tol_cg=1e-9, tol_cg_min=1e-9)
    opt, ora = make_pair(pkg, arrays, o)     # tight CG tolerance: both sides follow the exact Newton directions
    g, s = step_both(pkg, opt, ora, 3)
    assert abs(g.cg_iter_tot - s.cg_iter_tot) <= max(4, 0.2 * s.cg_iter_tot)
    assert abs(g.DIMACS_error - s.DIMACS_error) <= 1e-4 * max(1.0, s.DIMACS_error)
    assert abs(g.primal_obj - s.primal_obj) <= 1e-4 * (1 + abs(s.primal_obj))
    g.close()


def test_iteration_limit_status(pkg):
    """src/Solvers.jl:451-456: iter > maxit sets status 4 (the step still runs)."""
    arrays = pkg.problems.theta_torus(3, 4)
    opt = pkg.Optimizer()
    for k, v in dict(kit=0, initpoint=1, verb=0, maxit=2).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    opt.optimize()
    assert opt.solver.status == 4 and opt.solver.iter == 3 and opt.termination_status() == "ITERATION_LIMIT"
    opt.solver.close()


def test_maxG11_rank_one_end_to_end(pkg, golden_dir):
    """SDPLIB maxG11 (m = n_var = 800) through the rank-one Schur path (datarank = -1): SDPLIB optimum 629.1648, oracle
    iteration count +-1 (examples/solve_sdpa.jl:30 suggests exactly this instance for datarank = -1)."""
    z, arrays = golden(golden_dir, "maxG11")
    opt = pkg.Optimizer()
    for k, v in dict(OPTS_SDPA, datarank=-1).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    opt.optimize()
    s = opt.solver
    assert s.status == 1
    assert abs(s.iter - int(z["oracle_iters"])) <= 1
    assert abs(s.primal_obj - 629.1648) <= 1e-6 * 629.1648 * 10
    assert abs(s.primal_obj - float(z["oracle_obj"])) <= 1e-6 * (1 + abs(float(z["oracle_obj"])))
    s.close()


def test_theta_torus_full_size_known_optimum(pkg):
    """configs[2] at full size (m = 801, n_var = 2401, kit = 1, H_alpha): the Lovasz theta number of an even x even torus
    (bipartite, vertex transitive) is N/2 = 400 -- a size-independent known answer (SDPLIB thetaG11 = 400.00)."""
    c = pkg.problems.CONFIGS["C3"]
    opt = pkg.Optimizer()
    for k, v in dict(c["options"], verb=0).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*c["gen"]()))
    opt.optimize()
    s = opt.solver
    assert s.status == 1
    assert abs(s.primal_obj - 400.0) <= 400.0 * 2e-5            # eDIMACS = 1e-5
    assert abs(s.dual_obj - 400.0) <= 400.0 * 1e-4
    assert s.iter <= 40 and s.cg_iter_tot > 0
    s.close()


def test_large_blocks_properties_C2_quarter(pkg):
    """size-independent properties on a 25 x 50 max-cut instance (m = 1250, multi-pair block-Jacobi + Lanczos paths):
    W S W = X, G' S G = D, G Gi = I, H symmetric positive definite with H = W.^2 for F_k = e_k e_k'."""
    from loraine_jl_b200 import solver as S
    arrays = pkg.problems.maxcut_torus(25, 50, 5000)
    opt = pkg.Optimizer()
    for k, v in dict(kit=0, datarank=-1, initpoint=1, verb=0).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    g = opt.solver
    S.setup_solver(g, opt.halpha); S.initial_point(g)
    for _ in range(2):
        S.myIPstep(g, opt.halpha); g.itertime = 0.0; S.check_convergence(g)
    S.find_mu(g); S.prepare_W(g)
    W, G, Gi, D = g.get_array("W", 0), g.get_array("G", 0), g.get_array("GI", 0), g.get_array("D", 0)
    y, X, _ = S.get_solution(g)
    import ctypes as C
    m = g.model.msizes[0]
    Sm = [np.zeros((m, m), order="F")]
    Sp = (C.POINTER(C.c_double) * 1)(Sm[0].ctypes.data_as(C.POINTER(C.c_double)))
    g._call("lrn_get_slack", Sp, None)
    assert relerr(W @ Sm[0] @ W, X[0]) <= 1e-10
    assert relerr(G.T @ Sm[0] @ G, np.diag(D)) <= 1e-10
    assert np.linalg.norm(G @ Gi - np.eye(m)) <= 1e-9 * m
    g._call("lrn_residuals"); g._call("lrn_schur_assemble")
    H = g.get_array("H")
    assert relerr(H, W ** 2) <= 1e-11                          # b_k = e_k  =>  H = (W).^2
    assert g._call("lrn_schur_factor") == 0
    g.close()


def test_full_size_C2_properties(pkg):
    """BASELINE configs[1] at full size (max-cut n = m = 5000, datarank = -1): size-independent properties of one iteration,
    the heavy products checked against cuBLAS FP64 through torch (an implementation independent of this library):
    W S W = X, G' S G = D, H = W.^2 (F_k = e_k e_k'), L L' = H, H dely = rhs, and the predictor step keeps X, S positive
    definite (alpha, beta in (0, 1])."""
    import ctypes as C
    import torch
    from loraine_jl_b200 import solver as S
    cfg = pkg.problems.CONFIGS["C2"]
    opt = pkg.Optimizer()
    for k, v in dict(cfg["options"], verb=0).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*cfg["gen"]()))
    g = opt.solver
    S.setup_solver(g, opt.halpha); S.initial_point(g)
    for _ in range(2):
        S.myIPstep(g, opt.halpha); g.itertime = 0.0; S.check_convergence(g)
    g.iter += 1
    S.find_mu(g); S.prepare_W(g)
    m = g.model.msizes[0]
    assert m == 5000 and g.model.n == 5000
    dev = torch.device("cuda")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rel = lambda a, b: float(torch.linalg.norm(a - b) / torch.linalg.norm(b))
    W, G, D = T(g.get_array("W", 0)), T(g.get_array("G", 0)), T(g.get_array("D", 0))
    y, X, _ = S.get_solution(g)
    Sm = [np.zeros((m, m), order="F")]
    Sp = (C.POINTER(C.c_double) * 1)(Sm[0].ctypes.data_as(C.POINTER(C.c_double)))
    g._call("lrn_get_slack", Sp, None)
    Xt, St = T(X[0]), T(Sm[0])
    assert rel(W @ St @ W, Xt) <= 1e-10
    assert rel(G.T @ St @ G, torch.diag(D)) <= 1e-10
    S.predictor(g, opt.halpha)                                  # residuals, assembly, RHS, factorisation, solve, find_step
    H = T(g.get_array("H"))
    Hl = torch.tril(H)
    Hs = Hl + torch.tril(H, -1).T                               # the library fills (at least) the lower triangle
    assert rel(Hs, W * W) <= 1e-11                              # b_k = e_k  =>  H = W.^2
    L = torch.tril(T(g.get_array("L")))
    assert rel(L @ L.T, Hs) <= 1e-12
    dely, rhs = T(g.get_array("DELY")), T(g.get_array("RHS"))
    assert rel(Hs @ dely, rhs) <= 1e-9
    assert 0.0 < g.alpha[0] <= 1.0 and 0.0 < g.beta[0] <= 1.0
    g.close()


def schur_entry_ref(AA_csr, W, j, k):
    """H[j,k] = tr(A_j W A_k W) = sum_{(p,q) in A_j} sum_{(r,c) in A_k} a_pq a_rc W[q,r] W[c,p]  (src/makeBBBB.jl:39-64),
    A_j = mat(row j of AA) with vec index p + q*m."""
    m = W.shape[0]
    rj, rk = AA_csr.getrow(int(j)), AA_csr.getrow(int(k))
    pj, qj, vj = rj.indices % m, rj.indices // m, rj.data
    pk, qk, vk = rk.indices % m, rk.indices // m, rk.data
    # sum_{e in j} sum_{f in k} vj_e vk_f W[qj_e, pk_f] W[qk_f, pj_e]
    return float(np.einsum("e,f,ef,ef->", vj, vk, W[np.ix_(qj, pk)], W[np.ix_(pj, qk)]))


def test_full_size_C5_properties(pkg):
    """BASELINE configs[4] at full size (n_var = 40 000, one block m = 1000, general sparse assembly + 12.8 GB Cholesky):
    sampled entries of H against the defining formula tr(A_j W A_k W) evaluated in NumPy from the downloaded W,
    L L' = H against cuBLAS FP64 (torch), and H dely = rhs."""
    import torch
    from loraine_jl_b200 import solver as S
    cfg = pkg.problems.CONFIGS["C5"]
    opt = pkg.Optimizer()
    for k, v in dict(cfg["options"], verb=0).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*cfg["gen"]()))
    g = opt.solver
    S.setup_solver(g, opt.halpha); S.initial_point(g)
    S.myIPstep(g, opt.halpha); g.itertime = 0.0; S.check_convergence(g)
    g.iter += 1
    S.find_mu(g); S.prepare_W(g)
    S.predictor(g, opt.halpha)
    md = g.model
    n, m = md.n, md.msizes[0]
    assert n == 40000 and m == 1000
    W = g.get_array("W", 0)
    H = g.get_array("H")
    AA = md.AA[0].tocsr()
    rng = np.random.default_rng(0)
    js, ks = rng.integers(0, n, 200), rng.integers(0, n, 200)
    ref = np.array([schur_entry_ref(AA, W, j, k) for j, k in zip(js, ks)])
    got = H[js, ks]
    assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    # north_star tolerance on whole column panels: relative Frobenius error <= 1e-11 against the plain-C restatement of the
    # reference's pair formula (oracle/schur_pairs.c, pinned to the NumPy oracle in tests/test_oracle_golden.py) for the same W:
    # the first panel (all 40000 rows), one in the middle and the last one
    from oracle import c_oracle
    for k0 in (0, 19968, n - 512):
        ref_panel = c_oracle.schur_pairs_lower(md.AA[0], m, W, cols=(k0, k0 + 512))
        got_panel = np.tril(H[:, k0:k0 + 512], -k0)
        assert relerr(got_panel, ref_panel) <= 1e-11, k0
        del ref_panel, got_panel
    dev = torch.device("cuda")
    Ht = torch.from_numpy(H).to(dev)
    Lt = torch.tril(torch.from_numpy(g.get_array("L")).to(dev))
    R = Lt @ Lt.T
    R -= Ht
    assert float(torch.linalg.norm(R) / torch.linalg.norm(Ht)) <= 1e-12
    del R, Lt
    dely = torch.from_numpy(g.get_array("DELY")).to(dev)
    rhs = torch.from_numpy(g.get_array("RHS")).to(dev)
    assert float(torch.linalg.norm(Ht @ dely - rhs) / torch.linalg.norm(rhs)) <= 1e-9
    g.close()


@pytest.mark.parametrize("name,extra", [("tru9", {}), ("vib9", {}), ("thetaG11", dict(kit=1))])
def test_end_to_end_reference_data_large(pkg, golden_dir, name, extra):
    """The larger instances shipped in the reference's examples/data (tru9 / vib9: PSD blocks + a 6480-entry LP block,
    n_var = 3240; thetaG11: m = 801, n_var = 2401, the kit = 1 use case) with the options of examples/solve_sdpa.jl:
    objective and iteration count against the oracle's solve stored in the fixture (thetaG11 also: SDPLIB optimum 400)."""
    z, arrays = golden(golden_dir, name)
    opt = pkg.Optimizer()
    for k, v in dict(OPTS_SDPA, **extra).items():
        opt.set_attribute(k, v)
    opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
    opt.optimize()
    s = opt.solver
    assert s.status == 1
    ref_obj, ref_it = float(z["oracle_obj"]), int(z["oracle_iters"])
    assert abs(s.primal_obj - ref_obj) <= 2e-6 * (1 + abs(ref_obj))
    assert abs(s.iter - ref_it) <= 1
    if name == "thetaG11":
        assert abs(s.primal_obj - 400.0) <= 1e-3
        assert abs(s.cg_iter_tot - int(z["oracle_cg_iters"])) <= 0.2 * int(z["oracle_cg_iters"])
    s.close()
