"""GPU parity tests added in round 2: the sharded (multi-GPU) Schur path checked on ONE GPU, the distributed Cholesky with a
world-1 communicator, configs[3] at full size against the oracle, the H_alpha variants (erank > 1, aamat 0/1/3), direct
right-hand-side / residual comparisons and the in-process multi-GPU handle."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_solver import OPTS_SDPA, golden, make_pair, relerr, step_both

pytestmark = pytest.mark.gpu
PD = C.POINTER(C.c_double)


def _iterate(S, g):
    md = g.model
    y, X, xl = S.get_solution(g)
    Sm = [np.zeros((m, m), order="F") for m in md.msizes]
    sl = np.zeros(md.nlin)
    Sp = (PD * max(1, md.nlmi))(*[x.ctypes.data_as(PD) for x in Sm])
    g._call("lrn_get_slack", Sp, sl.ctypes.data_as(PD) if md.nlin else None)
    return [x.copy() for x in X], Sm, y.copy(), xl.copy(), sl


def _assembled_H(S, g):
    g.iter += 1
    S.find_mu(g); S.prepare_W(g)
    g._call("lrn_residuals"); g._call("lrn_schur_assemble")
    return g.get_array("H")


@pytest.mark.parametrize("case", ["control1", "tru3", "maxcut-rank1", "c5-mini"])
def test_row_block_shards_add_up_to_the_full_schur_matrix(pkg, golden_dir, case):
    """Multi-GPU assembly on ONE GPU: with the (rank, world) ownership set through the test hook (no communicator), every
    rank's lrn_schur_assemble must fill exactly its own row blocks of the lower triangle, the shards must be disjoint and
    their sum must be the single-GPU matrix (general path, LP term, rank-one path)."""
    from loraine_jl_b200 import solver as S
    if case == "maxcut-rank1":
        arrays, opts, br = pkg.problems.maxcut_torus(8, 12, 96), dict(OPTS_SDPA, datarank=-1), 32
    elif case == "c5-mini":
        arrays, opts, br = pkg.problems.large_schur(30, 200, 40000), dict(OPTS_SDPA), 64
    else:
        arrays, opts, br = golden(golden_dir, case)[1], dict(OPTS_SDPA), 8
    opt, ora = make_pair(pkg, arrays, opts)
    g0, _ = step_both(pkg, opt, ora, 2)
    it = _iterate(S, g0)
    H0 = np.tril(_assembled_H(S, g0))
    n = g0.model.n
    for world in (2, 3):
        total = np.zeros_like(H0)
        filled = np.zeros(H0.shape, dtype=np.int32)
        for rank in range(world):
            o2, _ = make_pair(pkg, arrays, opts)
            g = o2.solver
            S.setup_solver(g, o2.halpha)
            S.initial_point(g)
            S.set_iterate(g, *it)
            assert g.lib.lrn_dbg_set_shard(g.h, rank, world, br) == 0
            Hr = np.tril(_assembled_H(S, g))
            rows = np.nonzero(np.abs(Hr).sum(axis=1))[0]
            assert all(pkg.dist.row_owner(int(r), br, world) == rank for r in rows), (case, world, rank)
            # a sharded handle without a communicator must refuse to factor instead of silently factoring a shard
            with pytest.raises(Exception):
                g._call("lrn_schur_factor")
            total += Hr
            filled += (Hr != 0)
            g.close()
        assert filled.max() <= 1
        assert relerr(total, H0) <= 1e-13, (case, world)
        assert n == total.shape[0]
    g0.close()


def _nccl_world1(pkg, g):
    import torch  # noqa: F401  (loads the bundled libnccl into the process)
    buf = C.create_string_buffer(128)
    assert g.lib.lrn_dist_unique_id(buf) == 0
    g._call("lrn_dist_init", 0, 1, buf)


@pytest.mark.parametrize("shape", [(60, 700), (60, 1100)])
def test_distributed_cholesky_with_a_world1_communicator(pkg, shape):
    """cholesky_dist (row-block-cyclic: diagonal-block inverse broadcast, batched row solves, all-gather slots, strided-batch
    staircase updates) run with a one-rank NCCL communicator against the single-GPU look-ahead factorisation: same factor,
    same dely.  n_var = 700 / 1100 with 128-row blocks: several blocks and a partial last block."""
    from loraine_jl_b200 import solver as S
    arrays = pkg.problems.large_schur(shape[0], shape[1], 40000)
    opts = dict(kit=0, datarank=0, initpoint=1, verb=0, eDIMACS=1e-6)
    o1, _ = make_pair(pkg, arrays, opts)
    o2, _ = make_pair(pkg, arrays, opts)
    g1, g2 = o1.solver, o2.solver
    for g, o in ((g1, o1), (g2, o2)):
        S.setup_solver(g, o.halpha)
    _nccl_world1(pkg, g2)
    for g, o in ((g1, o1), (g2, o2)):
        S.initial_point(g)
        for _ in range(2):
            S.myIPstep(g, o.halpha); g.itertime = 0.0; S.check_convergence(g)
        g.iter += 1
        S.find_mu(g); S.prepare_W(g)
        g._call("lrn_residuals"); g._call("lrn_schur_assemble"); g._call("lrn_rhs_predictor")
        assert g._call("lrn_schur_factor") == 0
        g._call("lrn_schur_solve", 3)
    assert relerr(g2.get_array("H"), g1.get_array("H")) <= 1e-13
    L1, L2 = np.tril(g1.get_array("L")), np.tril(g2.get_array("L"))
    assert relerr(L2, L1) <= 1e-12
    H = g1.get_array("H")
    assert relerr(L2 @ L2.T, H) <= 1e-13
    assert relerr(g2.get_array("DELY"), g1.get_array("DELY")) <= 1e-10
    # a non-positive-definite matrix must report the same LAPACK-style pivot index on both paths
    for g in (g1, g2):
        g._call("lrn_schur_shift", -1e6)
    assert g1._call("lrn_schur_factor", allow_positive=True) == g2._call("lrn_schur_factor", allow_positive=True) > 0
    g1.close(); g2.close()


def test_full_size_C4_against_the_oracle(pkg):
    """configs[3] at FULL size (50 PSD blocks of side 200 + 2000 LP rows, n_var = 10000): the second IP iteration phase by
    phase against the oracle -- Schur matrix (block sum + LP term) to 1e-11 for the same W, dely, step lengths, DIMACS."""
    from oracle import loraine_oracle as lo
    from loraine_jl_b200 import solver as S
    cfg = pkg.problems.CONFIGS["C4"]
    opt, ora = make_pair(pkg, cfg["gen"](), dict(cfg["options"], verb=0))
    ora[1].lean = True                       # C restatement of the sparse assembly + in-place dpotrf (same arithmetic)
    g, s = step_both(pkg, opt, ora, 1)
    assert g.model.n == 10000 and g.model.nlmi == 50 and g.model.nlin == 2000
    for mod, st in ((S, g), (lo, s)):
        st.iter += 1
        st.cg_iter_pre = st.cg_iter_cor = 0
        mod.find_mu(st); mod.prepare_W(st)
    assert abs(g.mu - s.mu) <= 1e-9 * abs(s.mu)
    for i in (0, 17, 49):
        assert relerr(g.get_array("W", i), s.W[i]) <= 1e-9
    S.predictor(g, opt.halpha)
    H = np.tril(g.get_array("H"))
    keepW = s.W
    s.W = [g.get_array("W", i) for i in range(50)]
    _, _, _, xl, sl = _iterate(S, g)
    keep = s.X_lin, s.S_lin_inv
    s.X_lin, s.S_lin_inv = xl, 1.0 / sl
    Ho = np.tril(lo._assemble_lean(s))
    s.W = keepW
    s.X_lin, s.S_lin_inv = keep
    assert relerr(H, Ho) <= 1e-11
    del H, Ho
    lo.predictor(s, ora[2])
    assert np.linalg.norm(g.get_array("RP") - s.Rp) <= 1e-9 * (1.0 + np.linalg.norm(s.model.b))
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-7
    assert np.allclose(g.alpha, s.alpha, rtol=1e-6) and np.allclose(g.beta, s.beta, rtol=1e-6)
    assert abs(g.alpha_lin - s.alpha_lin) <= 1e-6 and abs(g.beta_lin - s.beta_lin) <= 1e-6
    assert abs(S.sigma_update(g) - lo.sigma_update(s)) <= 1e-6
    S.corrector(g, opt.halpha)
    lo.corrector(s, ora[2])
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-6
    S.check_convergence(g); lo.check_convergence(s)
    assert abs(g.DIMACS_error - s.DIMACS_error) <= 1e-6 * max(1.0, s.DIMACS_error)
    g.close()


@pytest.mark.parametrize("gen,args", [("multiblock_lp", (3, 12, 10, 7)), ("maxcut_torus", (6, 8, 3)), ("multiblock_lp", (3, 4, 40, 7))])
@pytest.mark.parametrize("erank,aamat", [(1, 0), (1, 1), (1, 3), (2, 2), (3, 2), (2, 0)])
def test_H_alpha_variants_match_the_oracle(pkg, gen, args, erank, aamat):
    """Prec_for_CG_tilS_prep / MyM (src/Solvers.jl:674-904) for every aamat (tau rule :646-655, identity term dropped for
    aamat = 3 :715-739) and for erank > 1 (the slow t = AA kron(U, Z) formula :752-768), with and without an LP block:
    M^-1 x against the oracle's functor built from the SAME W (the device's).  aamat = 3 drops tau^2 I, so AAAATtau is
    C_lin diag(x./s) C_lin' alone: only defined when the LP block has at least n_var independent rows (third instance:
    12 variables, 40 LP rows); the instances use random weights (no repeated eigenvalues of W: the erank leading
    eigenvectors must be unique)."""
    from oracle import loraine_oracle as lo
    from loraine_jl_b200 import solver as S
    full_rank_lp = (gen == "multiblock_lp" and args[2] >= 3 * args[0] * args[1])
    if aamat == 3 and not full_rank_lp:
        pytest.skip("aamat = 3 needs an LP block of full row rank (AAAATtau = C_lin D C_lin' must be invertible)")
    if full_rank_lp and erank >= 3:
        pytest.skip("erank must stay below m - 1 = 3")
    arrays = getattr(pkg.problems, gen)(*args)
    o = dict(kit=1, preconditioner=1, erank=erank, aamat=aamat, initpoint=1, verb=0, eDIMACS=1e-6)
    opt, ora = make_pair(pkg, arrays, dict(o, aamat=2 if aamat == 3 else aamat))     # two ordinary iterations first
    g, s = step_both(pkg, opt, ora, 2)
    g._call("lrn_set_option", b"aamat", float(aamat))
    s.aamat = aamat
    for mod, st in ((S, g), (lo, s)):
        st.iter += 1; st.cg_iter_pre = st.cg_iter_cor = 0
        mod.find_mu(st); mod.prepare_W(st)
    g._call("lrn_residuals"); g._call("lrn_rhs_predictor"); g._call("lrn_prec_prepare", 1)
    ha = ora[2]
    # the oracle's factors from the DEVICE's iterate (W, x_lin ./ s_lin): differences are then the algorithm's alone
    s.W = [g.get_array("W", i) for i in range(s.model.nlmi)]
    if s.model.nlin:
        _, _, _, xl, sl = _iterate(S, g)
        s.X_lin, s.S_lin_inv = xl, 1.0 / sl
    lo.Prec_for_CG_tilS_prep(s, ha)
    rng = np.random.default_rng(erank * 10 + aamat)
    dpx = lambda a: a.ctypes.data_as(PD)
    for _ in range(2):
        x = rng.standard_normal(s.model.n)
        out = np.zeros_like(x)
        g._call("lrn_apply_operator", 1, dpx(x), dpx(out))
        assert relerr(out, lo.MyM(s, ha)(x)) <= 1e-6, (erank, aamat)
    g.close()


@pytest.mark.parametrize("name,datarank", [("control1", 0), ("vib3", 0), ("maxcut", -1)])
def test_right_hand_sides_and_residuals_directly(pkg, golden_dir, name, datarank):
    """LRN_ARR_RP / LRN_ARR_RD / LRN_ARR_RHS (include/loraine_b200.h) against the oracle's Rp, Rd_i and both right-hand sides
    (makeRHS src/makeBBBB.jl:221-228; corrector RHS src/predictor_corrector.jl:183-192) on the same iterate."""
    from oracle import loraine_oracle as lo
    from loraine_jl_b200 import solver as S
    arrays = pkg.problems.maxcut_torus(8, 12, 96) if name == "maxcut" else golden(golden_dir, name)[1]
    opt, ora = make_pair(pkg, arrays, dict(OPTS_SDPA, datarank=datarank))
    g, s = step_both(pkg, opt, ora, 2)
    S.set_iterate(g, *_iterate(S, g))            # round trip through the ABI (upload what was downloaded)
    got = {}
    s.hooks["dely_pred"] = lambda s_, h, d: got.__setitem__("pred", h.copy())
    s.hooks["rhs_corr"] = lambda s_, h: got.__setitem__("corr", h.copy())
    for mod, st, ha in ((S, g, opt.halpha), (lo, s, ora[2])):
        st.iter += 1
        st.cg_iter_pre = st.cg_iter_cor = 0
        mod.find_mu(st); mod.prepare_W(st); mod.predictor(st, ha)
    nb = 1.0 + np.linalg.norm(s.model.b)           # Rp tends to zero (primal feasibility): error relative to 1 + ||b|| like err1
    assert np.linalg.norm(g.get_array("RP") - s.Rp) <= 1e-9 * nb
    for i in range(s.model.nlmi):
        assert np.linalg.norm(g.get_array("RD", i) - s.Rd[i]) <= 1e-9 * (1.0 + np.linalg.norm(s.model.C[i].toarray()))
    assert relerr(g.get_array("RHS"), got["pred"]) <= 1e-8
    assert abs(S.sigma_update(g) - lo.sigma_update(s)) <= 1e-6
    S.corrector(g, opt.halpha); lo.corrector(s, ora[2])
    assert relerr(g.get_array("RHS"), got["corr"]) <= 1e-7
    g.close()


def test_in_process_multi_gpu_handle(pkg):
    """lrn_create_multi: one host thread, N devices behind ONE handle.  With one visible GPU the call must
    degrade to the single-device handle; with >= 2 it must solve configs[4]-mini to the single-GPU answer."""
    import torch
    from loraine_jl_b200 import solver as S
    ndev = torch.cuda.device_count()
    cfg = pkg.problems.CONFIGS["C5-mini"]
    arrays = pkg.problems.large_schur(60, 700, 40000)
    res = []
    for ngpus in sorted({1, min(2, ndev), ndev}):
        opt = pkg.Optimizer()
        for k, v in dict(cfg["options"], verb=0, ngpus=ngpus).items():
            opt.set_attribute(k, v)
        opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
        opt.optimize()
        s = opt.solver
        assert s.status == 1
        res.append((s.iter, s.primal_obj, s.y.copy()))
        s.close()
    for it, obj, y in res[1:]:
        assert it == res[0][0] and abs(obj - res[0][1]) <= 1e-9 * (1 + abs(obj))
        assert relerr(y, res[0][2]) <= 1e-7


@pytest.mark.parametrize("name,datarank", [("theta1", 0), ("control1", 0), ("tru3", 0), ("vib3", 0), ("maxG11-mini", -1)])
def test_model_preparation_inside_the_library(pkg, golden_dir, name, datarank, tmp_path):
    """lrn_create_from_triplets / lrn_load_sdpa / lrn_initial_point (C++: prep_AA!, prep_B, prep_sparse!,
    find_initial!) against the Python host preparation (model.py + lrn_set_block_* + lrn_finalize): identical Schur matrix on
    the first iterations and the same solve; the .dat-s route gives the same handle as the triplet route."""
    from loraine_jl_b200 import solver as S, model as M, _lib
    arrays = pkg.problems.maxcut_torus(8, 12, 96) if name == "maxG11-mini" else golden(golden_dir, name)[1]
    opts = dict(OPTS_SDPA, datarank=datarank)
    runs = []
    for native in (False, True):
        opt = pkg.Optimizer()
        for k, v in opts.items():
            opt.set_attribute(k, v)
        if native:
            opt.load_sdpa(*arrays)
        else:
            opt.copy_to(pkg.raw_from_sdpa_arrays(*arrays))
        g = opt.solver
        S.setup_solver(g, opt.halpha); S.initial_point(g)
        assert bool(getattr(g, "native_model", False)) == native
        Hs = []
        for _ in range(2):
            Hs.append(_assembled_H(S, g).copy())
            g.iter -= 1
            S.myIPstep(g, opt.halpha); g.itertime = 0.0; S.check_convergence(g)
        S.solve(g, opt.halpha)
        runs.append((Hs, g.iter, g.primal_obj, g.y.copy(), g.status))
        g.close()
    (H0, it0, ob0, y0, st0), (H1, it1, ob1, y1, st1) = runs
    for a, b in zip(H0, H1):
        assert relerr(b, a) <= 1e-13
    assert st0 == st1 == 1 and it0 == it1
    assert abs(ob0 - ob1) <= 1e-9 * (1 + abs(ob0)) and relerr(y1, y0) <= 1e-8
    # the file route: write the instance as .dat-s, load it with the library's own reader, solve through the C ABI
    path = str(tmp_path / (name + ".dat-s"))
    M.write_sdpa(path, *arrays)
    L = _lib.lib()
    o = _lib.lrn_options_t()
    L.lrn_default_options(C.byref(o))
    o.kit, o.datarank, o.datasparsity = 0, datarank, 8
    h = C.c_void_p()
    assert L.lrn_load_sdpa(C.byref(h), path.encode(), C.byref(o), 1) == 0
    assert L.lrn_initial_point(h, 1) == 0
    mu = C.c_double()
    assert L.lrn_find_mu(h, C.byref(mu)) == 0 and mu.value > 0
    st4 = C.c_int32()
    assert L.lrn_prepare_W(h, C.byref(st4)) == 0 and st4.value == 0
    assert L.lrn_residuals(h) == 0 and L.lrn_schur_assemble(h) == 0
    n = int(arrays[0])
    H = np.zeros((n, n), order="F")
    assert L.lrn_get_array(h, 1, 0, H.ctypes.data_as(PD)) == 0
    assert relerr(H, H0[0]) <= 1e-13
    L.lrn_destroy(h)


@pytest.mark.parametrize("name,want", [("theta1", 23.0), ("tru3", None), ("control1", 17.78463)])
def test_plain_c_host_solves_through_the_header(pkg, golden_dir, tmp_path, name, want):
    """tests/c_abi_host.c (C, include/loraine_b200.h only: lrn_load_sdpa, lrn_initial_point and the per-iteration entry
    points) solves the reference's own instances: theta1 -> 23 (examples/solve_sdpa.jl:61), control1 -> SDPLIB 17.78463,
    tru3 (LP block) -> the fixture's objective, in the fixture's number of iterations +-1."""
    import subprocess
    from test_host_cpu import _build_c_host
    from loraine_jl_b200 import model as M
    z, arrays = golden(golden_dir, name)
    path = str(tmp_path / (name + ".dat-s"))
    M.write_sdpa(path, *arrays)
    exe = _build_c_host(tmp_path)
    r = subprocess.run([exe, path, "1e-6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    tok = r.stdout.split()
    obj, iters, status = float(tok[1]), int(tok[3]), int(tok[5])
    assert status == 1
    ref = float(z["oracle_obj"])
    assert abs(obj - ref) <= 1e-6 * (1 + abs(ref))
    assert abs(iters - int(z["oracle_iters"])) <= 1
    if want is not None:
        assert abs(obj - want) <= 2e-6 * want
    assert int(tok[-1]) > 0                     # kernels were launched by the library in that process


@pytest.mark.parametrize("case", ["c5-mid", "maxcut-general", "tru3", "vib3", "theta"])
def test_staged_pair_kernel_matches_gather_kernel_and_oracle(pkg, golden_dir, case):
    """Sparse-pair Schur term (F3 formula, src/makeBBBB.jl:139-213): the shared-memory staged kernel (csrc/pairs.cu: rows of W
    staged per group of 8 constraints, lower-triangle entry lists with doubled weights, bucketed by entry count) against the
    one-thread-per-pair gather kernel (1e-13) and against the plain-C restatement of the reference formula (1e-11), including
    2-to-25-entry matrices, single diagonal entries (max-cut with datarank = 0), several blocks and an LP term."""
    from oracle import c_oracle, loraine_oracle as lo
    from loraine_jl_b200 import solver as S
    if case == "c5-mid":
        arrays = pkg.problems.large_schur(120, 3000, 40000)
    elif case == "maxcut-general":
        arrays = pkg.problems.maxcut_torus(12, 16, 11)
    elif case == "theta":
        arrays = pkg.problems.theta_torus(6, 8)
    else:
        arrays = golden(golden_dir, case)[1]
    opt, ora = make_pair(pkg, arrays, dict(OPTS_SDPA))
    g, s = step_both(pkg, opt, ora, 2)
    g.iter += 1
    S.find_mu(g); S.prepare_W(g); g._call("lrn_residuals")
    Hs = {}
    for mode in (1.0, 0.0):
        g._call("lrn_set_option", b"pair_kernel", mode)
        g._call("lrn_schur_assemble")
        Hs[mode] = np.tril(g.get_array("H"))
    assert relerr(Hs[1.0], Hs[0.0]) <= 1e-13
    md = s.model
    Href = np.zeros((md.n, md.n), order="F")
    for i in range(md.nlmi):
        c_oracle.schur_pairs_lower(md.AA[i], md.msizes[i], g.get_array("W", i), H=Href, accumulate=(i > 0))
    if md.nlin:
        _, _, _, xl, sl = _iterate(S, g)
        Href += np.tril(lo.lp_schur(md, xl / sl))
    assert relerr(Hs[1.0], np.tril(Href)) <= 1e-11
    g.close()


def test_lanczos_non_convergence_falls_back_to_guaranteed_bounds(pkg):
    """an unconverged Ritz value lies above lambda_min, so the step length would be overestimated silently.  With the
    Krylov budget forced down to 6 vectors (lrn_set_option lanczos_kmax) every Lanczos run of find_step fails to converge and
    the Cholesky bisection fallback must still deliver the oracle's step lengths (exact eigmin) on an m = 480 block (blocks
    up to 384 use the tridiagonalisation kernel instead of Lanczos)."""
    from loraine_jl_b200 import solver as S
    arrays = pkg.problems.maxcut_torus(20, 24, 11)
    opt, ora = make_pair(pkg, arrays, dict(kit=0, datarank=-1, initpoint=1, verb=0))
    g, s = step_both(pkg, opt, ora, 3)
    lo = ora[0]
    g._call("lrn_set_option", b"lanczos_kmax", 6.0)
    for mod, st, ha in ((S, g, opt.halpha), (lo, s, ora[2])):
        st.iter += 1
        mod.find_mu(st); mod.prepare_W(st); mod.predictor(st, ha)
    assert g.stats()["lanczos_not_converged"] > 0
    assert np.allclose(g.alpha, s.alpha, rtol=1e-6) and np.allclose(g.beta, s.beta, rtol=1e-6)
    assert abs(S.sigma_update(g) - lo.sigma_update(s)) <= 1e-6
    S.corrector(g, opt.halpha); lo.corrector(s, ora[2])
    assert np.allclose(g.alpha, s.alpha, rtol=1e-6) and np.allclose(g.beta, s.beta, rtol=1e-6)
    assert relerr(g.get_array("DELY"), s.dely) <= 1e-6
    g.close()
