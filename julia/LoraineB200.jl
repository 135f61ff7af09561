# LoraineB200.jl -- `ccall` shims that put libloraine_b200.so behind the unchanged Loraine.Optimizer / MOI / JuMP surface.
#
# Usage (on a machine with Julia, Loraine.jl v0.2.5 and a B200):
#     using Loraine; include("LoraineB200.jl"); LoraineB200.enable!("/path/to/libloraine_b200.so"; ngpus = 1)
#     model = Model(Loraine.Optimizer); ... ; optimize!(model)          # unchanged user code
#
# How it hooks in (no method of the reference is overwritten; every method below is MORE SPECIFIC than the reference's and falls
# back to the reference's own method through `invoke` until `enable!` has been called):
#   * `Solvers.setup_solver(::MySolver{Float64}, ::Halpha)`  runs the reference body (src/Solvers.jl:363-446) via `invoke`,
#     then creates the device handle and uploads the prepared model (`attach!`);
#   * `Solvers.initial_point(::MySolver{Float64})`           runs the reference body (src/initial_point.jl:1-81), then uploads
#     X, S, y, X_lin, S_lin (`upload_iterate!`);
#   * `find_mu`, `prepare_W`, `predictor`, `sigma_update`, `corrector`, `find_step`, `myIPstep`, `check_convergence` for
#     `MySolver{Float64}` keep the reference's control flow (src/Solvers.jl:448-568, src/predictor_corrector.jl:5-364) with
#     every array expression replaced by one ccall;
#   * `Solvers.solve(::MySolver{Float64}, ::Halpha)`         runs the reference loop (src/Solvers.jl:304-361) via `invoke`,
#     then downloads y, X, X_lin so that the MOI getters (src/MOI_wrapper.jl:241-354) read valid host arrays;
#   * `Solvers.solve(::MySolver{T}, ::Halpha) where {T<:AbstractFloat}` (every other element type, e.g. Float64x2 of
#     examples/k.jl:8) throws an ArgumentError: there is no fallback.
#
# STATUS: UNTESTED UNDER JULIA.  Julia is not available in the build image (no network either), so this file has been
# written against the reference sources and include/loraine_b200.h and has never been executed.  The executed mirrors of
# the same control flow are loraine.jl_b200/solver.py (Python, all GPU tests) and tests/c_abi_host.c (plain C).
module LoraineB200

using Loraine
using SparseArrays, LinearAlgebra, Printf
using TimerOutputs
const S = Loraine.Solvers

const LIB = Ref{String}("libloraine_b200.so")
const NGPUS = Ref{Int32}(1)          # 1: one device; otherwise lrn_create_multi (one host thread, N devices, NCCL inside the library)
const ENABLED = Ref{Bool}(false)

struct Options              # mirrors lrn_options_t (include/loraine_b200.h)
    kit::Int32; datarank::Int32; preconditioner::Int32; erank::Int32; aamat::Int32; datasparsity::Int32
    schur_split::Int32; rank1_mode::Int32; svd_tol::Float64; lanczos_tol::Float64; device::Int32; reserved::Int32
end

mutable struct Device
    h::Ptr{Cvoid}
    function Device(h)
        d = new(h)
        finalizer(x -> (x.h != C_NULL && ccall((:lrn_destroy, LIB[]), Int32, (Ptr{Cvoid},), x.h); x.h = C_NULL), d)
        return d
    end
end
const DEVICES = IdDict{Any,Device}()       # MySolver => Device

lasterr(d) = unsafe_string(ccall((:lrn_last_error, LIB[]), Cstring, (Ptr{Cvoid},), d.h))
function check(d, rc, what)
    rc < 0 && error("loraine_b200: $what failed ($rc): $(lasterr(d))")
    rc > 0 && throw(LinearAlgebra.PosDefException(rc))       # keeps the reference's try/catch working (predictor :56-88)
    return rc
end

csc(A::SparseMatrixCSC{Float64,Int64}) = (A.colptr, A.rowval, A.nzval)

"Set the library path (and the number of GPUs one solve may use) and check that the library can be loaded."
function enable!(libpath::AbstractString = LIB[]; ngpus::Integer = 1)
    LIB[] = libpath
    NGPUS[] = Int32(ngpus)
    ccall((:lrn_kernel_launches, LIB[]), Int64, ())          # throws when the library cannot be loaded: no CPU fallback
    ENABLED[] = true
    return nothing
end

"Create the device handle and upload the prepared model (outputs of `_prepare_A`, src/model.jl:120-150)."
function attach!(solver::S.MySolver{Float64})
    ENABLED[] || error("LoraineB200.enable!(path) was not called")
    haskey(DEVICES, solver) && delete!(DEVICES, solver)          # a second solve on the same object: fresh handle
    md = solver.model
    opt = Ref(Options(solver.kit, solver.datarank, solver.preconditioner, solver.erank, solver.aamat, solver.datasparsity,
                      0, 0, 0.0, 0.0, -1, 0))
    hr = Ref{Ptr{Cvoid}}(C_NULL)
    ms = Int64.(md.msizes)
    rc = if NGPUS[] == 1
        ccall((:lrn_create, LIB[]), Int32, (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Int64, Ref{Options}),
              hr, md.n, md.nlmi, ms, md.nlin, opt)
    else
        ccall((:lrn_create_multi, LIB[]), Int32, (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Int64, Ref{Options}, Int32, Ptr{Int32}),
              hr, md.n, md.nlmi, ms, md.nlin, opt, NGPUS[], C_NULL)
    end
    rc == 0 || error("loraine_b200: lrn_create failed ($rc); there is no CPU fallback")
    d = Device(hr[])
    for i in 1:md.nlmi
        AAi = SparseMatrixCSC{Float64,Int64}(md.AA[i])     # n x m^2 (row k = vec(calA_k))
        GC.@preserve AAi check(d, ccall((:lrn_set_block_AA, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                        d.h, i - 1, csc(AAi)...), "lrn_set_block_AA")
        Ci = SparseMatrixCSC{Float64,Int64}(md.C[i])
        GC.@preserve Ci check(d, ccall((:lrn_set_block_C, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                       d.h, i - 1, csc(Ci)...), "lrn_set_block_C")
        if solver.datarank == -1 && !isempty(md.B)
            Bi = SparseMatrixCSC{Float64,Int64}(md.B[i])
            GC.@preserve Bi check(d, ccall((:lrn_set_block_B, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                           d.h, i - 1, csc(Bi)...), "lrn_set_block_B")
        end
    end
    if md.nlin > 0
        Cl = SparseMatrixCSC{Float64,Int64}(md.C_lin); dl = Vector{Float64}(vec(md.d_lin))
        GC.@preserve Cl dl check(d, ccall((:lrn_set_lin, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                                          d.h, csc(Cl)..., dl), "lrn_set_lin")
    end
    b = Vector{Float64}(vec(md.b))
    check(d, ccall((:lrn_set_b, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), d.h, b), "lrn_set_b")
    check(d, ccall((:lrn_finalize, LIB[]), Int32, (Ptr{Cvoid},), d.h), "lrn_finalize")
    DEVICES[solver] = d
    return d
end

dev(solver) = DEVICES[solver]

function upload_iterate!(solver)          # after initial_point (src/initial_point.jl)
    d = dev(solver); md = solver.model
    X = [Matrix{Float64}(x) for x in solver.X]; Sm = [Matrix{Float64}(x) for x in solver.S]
    Xp = Ptr{Float64}[pointer(x) for x in X]; Sp = Ptr{Float64}[pointer(x) for x in Sm]
    y = Vector{Float64}(vec(solver.y)); xl = Vector{Float64}(vec(solver.X_lin)); sl = Vector{Float64}(vec(solver.S_lin))
    GC.@preserve X Sm Xp Sp y xl sl check(d, ccall((:lrn_set_iterate, LIB[]), Int32,
        (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        d.h, Xp, Sp, y, md.nlin > 0 ? pointer(xl) : C_NULL, md.nlin > 0 ? pointer(sl) : C_NULL), "lrn_set_iterate")
end

function download_solution!(solver)       # MOI getters read solver.y / X / X_lin (src/MOI_wrapper.jl:315-354)
    d = dev(solver); md = solver.model
    y = zeros(md.n); X = [zeros(Int(m), Int(m)) for m in md.msizes]; xl = zeros(md.nlin)
    Xp = Ptr{Float64}[pointer(x) for x in X]
    GC.@preserve X Xp y xl check(d, ccall((:lrn_get_solution, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Ptr{Float64}}, Ptr{Float64}),
                                          d.h, y, Xp, md.nlin > 0 ? pointer(xl) : C_NULL), "lrn_get_solution")
    solver.y = y; solver.X = X; solver.X_lin = xl
end

# --- the `solver.cholBBBB` seam (src/predictor_corrector.jl:57,85,89-90,199) ------------------------------------------
struct DeviceFactor
    solver::Any
    is_cholesky_object::Bool   # false: plays L (normal path); true: plays Julia's `Cholesky` (after a regularised retry)
end
solve_reference_expression!(f::DeviceFactor) =   # dely = cholBBBB' \ (cholBBBB \ h)
    check(dev(f.solver), ccall((:lrn_schur_solve, LIB[]), Int32, (Ptr{Cvoid}, Int32), dev(f.solver).h,
                               f.is_cholesky_object ? 6 : 3), "lrn_schur_solve")

# (ccall needs a literal function name, hence one method per entry point)
call0(solver, ::Val{:lrn_residuals}) = check(dev(solver), ccall((:lrn_residuals, LIB[]), Int32, (Ptr{Cvoid},), dev(solver).h), "lrn_residuals")
call0(solver, ::Val{:lrn_schur_assemble}) = check(dev(solver), ccall((:lrn_schur_assemble, LIB[]), Int32, (Ptr{Cvoid},), dev(solver).h), "lrn_schur_assemble")
call0(solver, ::Val{:lrn_rhs_predictor}) = check(dev(solver), ccall((:lrn_rhs_predictor, LIB[]), Int32, (Ptr{Cvoid},), dev(solver).h), "lrn_rhs_predictor")
call0(solver, ::Val{:lrn_schur_factor}) = check(dev(solver), ccall((:lrn_schur_factor, LIB[]), Int32, (Ptr{Cvoid},), dev(solver).h), "lrn_schur_factor")

# =====================================================================================================================
#  methods of module Solvers for MySolver{Float64} (more specific than the reference's; nothing is overwritten)
# =====================================================================================================================
function S.setup_solver(solver::S.MySolver{Float64}, halpha::S.Halpha)
    invoke(S.setup_solver, Tuple{S.MySolver,S.Halpha}, solver, halpha)      # the reference body: kit / datarank fall-backs, host zeros
    ENABLED[] || return nothing                                             # not enabled: the reference's CPU path, untouched
    attach!(solver)
    return nothing
end

function S.initial_point(solver::S.MySolver{Float64})
    invoke(S.initial_point, Tuple{Any}, solver)                            # src/initial_point.jl:1-81 (a few norms of the model data)
    ENABLED[] || return nothing
    upload_iterate!(solver)
    return nothing
end

function S.find_mu(solver::S.MySolver{Float64})                            # src/Solvers.jl:480-494
    ENABLED[] || return invoke(S.find_mu, Tuple{Any}, solver)
    mu = Ref(0.0)
    check(dev(solver), ccall((:lrn_find_mu, LIB[]), Int32, (Ptr{Cvoid}, Ref{Float64}), dev(solver).h, mu), "lrn_find_mu")
    solver.mu = mu[]
end

function S.prepare_W(solver::S.MySolver{Float64})                          # src/prepare_W.jl:28-94
    ENABLED[] || return invoke(S.prepare_W, Tuple{S.MySolver}, solver)
    st4 = Ref{Int32}(0)
    @timeit solver.to "prep W SVD" begin                                   # section name of src/prepare_W.jl:37
        check(dev(solver), ccall((:lrn_prepare_W, LIB[]), Int32, (Ptr{Cvoid}, Ref{Int32}), dev(solver).h, st4), "lrn_prepare_W")
    end
    if st4[] != 0
        solver.verb > 0 && println("WARNING: X or S cannot be made positive definite, giving up")
        solver.status = 4
    end
end

function pcg(solver, kind)
    it = Ref{Int64}(0); code = Ref{Int32}(0)
    check(dev(solver), ccall((:lrn_pcg, LIB[]), Int32, (Ptr{Cvoid}, Float64, Int64, Int32, Ref{Int64}, Ref{Int32}),
                             dev(solver).h, Float64(solver.tol_cg), 10000, kind, it, code), "lrn_pcg")
    return it[]
end

function S.find_step(solver::S.MySolver{Float64})                          # src/predictor_corrector.jl:248-364
    ENABLED[] || return invoke(S.find_step, Tuple{S.MySolver}, solver)
    d = dev(solver); md = solver.model
    a = zeros(max(1, md.nlmi)); b = zeros(max(1, md.nlmi)); al = Ref(1.0); bl = Ref(1.0)
    check(d, ccall((:lrn_find_step, LIB[]), Int32,
                   (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Float64}),
                   d.h, solver.predict ? 1 : 0, Float64(solver.sigma), Float64(solver.mu), Float64(solver.tau), a, b, al, bl),
          "lrn_find_step")
    solver.alpha = a[1:md.nlmi]; solver.beta = b[1:md.nlmi]; solver.alpha_lin = al[]; solver.beta_lin = bl[]
end

function S.predictor(solver::S.MySolver{Float64}, halpha::S.Halpha)        # src/predictor_corrector.jl:5-146, control flow kept
    ENABLED[] || return invoke(S.predictor, Tuple{S.MySolver,S.Halpha}, solver, halpha)
    solver.predict = true
    call0(solver, Val(:lrn_residuals))                                     # :8-22
    if solver.kit == 0                                                     # :24-40 (section names of src/makeBBBB.jl:2,30)
        @timeit solver.to (solver.datarank == -1 ? "BBBB_rank1" : "BBBBs") call0(solver, Val(:lrn_schur_assemble))
    end
    call0(solver, Val(:lrn_rhs_predictor))                                 # :43-50
    if solver.kit == 0
        try
            call0(solver, Val(:lrn_schur_factor))                          # :57
            solver.cholBBBB = DeviceFactor(solver, false)
        catch err
            err isa LinearAlgebra.PosDefException || rethrow()
            solver.verb > 0 && println("Matrix H not positive definite, trying to regularize")
            icount = 0
            solver.regcount += 1
            if solver.regcount > 5
                solver.verb > 0 && println("WARNING: too many regularizations of H, giving up")
                solver.status = 3
                return
            end
            while true                                                    # while isposdef(BBBB) == false  (:73-84)
                check(dev(solver), ccall((:lrn_schur_shift, LIB[]), Int32, (Ptr{Cvoid}, Float64), dev(solver).h, 1e-4), "lrn_schur_shift")
                icount += 1
                rc = ccall((:lrn_schur_factor, LIB[]), Int32, (Ptr{Cvoid},), dev(solver).h)
                rc < 0 && error("loraine_b200: lrn_schur_factor failed: $(lasterr(dev(solver)))")
                rc == 0 && break
                if icount > 1000
                    solver.verb > 0 && println("WARNING: H cannot be made positive definite, giving up")
                    solver.status = 3
                    return
                end
            end
            solver.cholBBBB = DeviceFactor(solver, true)                  # `cholesky(BBBB)` object, :85
        end
        solve_reference_expression!(solver.cholBBBB)                      # :89-90
    else
        kind = solver.preconditioner == 0 ? 0 : solver.preconditioner == 1 ? 1 : 2
        solver.preconditioner == 3 && error("preconditioner 3 is undefined in the reference (src/predictor_corrector.jl:120-128)")
        # the hybrid switch of the reference loop mutates solver.aamat / solver.preconditioner on the host (src/Solvers.jl:339-347)
        check(dev(solver), ccall((:lrn_set_option, LIB[]), Int32, (Ptr{Cvoid}, Cstring, Float64), dev(solver).h, "aamat", Float64(solver.aamat)), "lrn_set_option")
        if kind != 0
            @timeit solver.to "prec" begin                                 # section name of src/Solvers.jl:676
                check(dev(solver), ccall((:lrn_prec_prepare, LIB[]), Int32, (Ptr{Cvoid}, Int32), dev(solver).h, kind), "lrn_prec_prepare")
            end
        end
        n = pcg(solver, kind)
        solver.cg_iter_pre += n; solver.cg_iter_tot += n
    end
    S.find_step(solver)
end

function S.sigma_update(solver::S.MySolver{Float64})                       # src/predictor_corrector.jl:148-179
    ENABLED[] || return invoke(S.sigma_update, Tuple{S.MySolver}, solver)
    md = solver.model
    step_pred = min(minimum([solver.alpha; solver.alpha_lin]), minimum([solver.beta; solver.beta_lin]))
    expon_used = solver.mu > 1e-6 ? (step_pred < 1 / sqrt(3) ? 1.0 : max(solver.expon, 3 * step_pred^2)) :
                 max(1, min(solver.expon, 3 * step_pred^2))
    tr = Ref(0.0); dl = Ref(0.0)
    check(dev(solver), ccall((:lrn_sigma_trace, LIB[]), Int32, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), dev(solver).h, tr, dl), "lrn_sigma_trace")
    if tr[] < 0
        solver.sigma = 0.8
    else
        tmp12 = ((md.nlmi > 0 ? tr[] : 0.0) + (md.nlin > 0 ? dl[] : 0.0)) / (sum(md.msizes) + md.nlin)
        solver.sigma = min(1.0, (tmp12 / Float64(solver.mu))^Float64(expon_used))
    end
    return solver.sigma
end

function S.corrector(solver::S.MySolver{Float64}, halpha)                  # src/predictor_corrector.jl:181-246
    ENABLED[] || return invoke(S.corrector, Tuple{Any,Any}, solver, halpha)
    solver.predict = false
    check(dev(solver), ccall((:lrn_rhs_corrector, LIB[]), Int32, (Ptr{Cvoid}, Float64, Float64), dev(solver).h,
                             Float64(solver.sigma), Float64(solver.mu)), "lrn_rhs_corrector")
    if solver.kit == 0
        solve_reference_expression!(solver.cholBBBB)                      # :199
    else
        kind = solver.preconditioner == 0 ? 0 : solver.preconditioner == 1 ? 1 : 2
        n = pcg(solver, kind)
        solver.cg_iter_cor += n; solver.cg_iter_tot += n
    end
    S.find_step(solver)
end

function S.myIPstep(solver::S.MySolver{Float64}, halpha::S.Halpha)         # src/Solvers.jl:448-478
    ENABLED[] || return invoke(S.myIPstep, Tuple{S.MySolver,S.Halpha}, solver, halpha)
    solver.iter += 1
    if solver.iter > solver.maxit
        solver.status = 4
        solver.verb > 0 && println("WARNING: Stopped by iteration limit (stopping status = 4)")
    end
    solver.cg_iter_pre = 0
    solver.cg_iter_cor = 0
    S.find_mu(solver)
    S.prepare_W(solver)
    @timeit solver.to "predictor" S.predictor(solver, halpha)
    # H could not be made positive definite (src/predictor_corrector.jl:66-70, :76-83): the reference finishes the iteration
    # with cholBBBB = I and leaves the loop; there is no factor on the device, so the iteration ends here with status 3
    solver.status == 3 && return
    S.sigma_update(solver)
    @timeit solver.to "corrector" S.corrector(solver, halpha)
end

function S.check_convergence(solver::S.MySolver{Float64})                  # src/Solvers.jl:496-568
    ENABLED[] || return invoke(S.check_convergence, Tuple{Any}, solver)
    md = solver.model
    err = zeros(6); by = Ref(0.0); trCX = Ref(0.0); dx = Ref(0.0)
    check(dev(solver), ccall((:lrn_dimacs, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}),
                             dev(solver).h, err, by, trCX, dx), "lrn_dimacs")
    solver.err1, solver.err2, solver.err3, solver.err4, solver.err5, solver.err6 = err
    DIMACS_error = md.nlmi > 0 ? err[1] + err[2] + err[3] + err[4] + abs(err[5]) + err[6] :
                                 err[2] + err[3] + err[4] + abs(err[5]) + err[6]
    solver.DIMACS_error = DIMACS_error
    obj = -by[] + md.b_const
    if solver.verb > 0 && solver.status == 0
        if solver.verb > 1
            if solver.kit == 0
                @printf("%3.0d %16.8e %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e %8.2f\n", solver.iter, obj, DIMACS_error, err[1], err[2], err[3], err[4], err[5], err[6], solver.itertime)
            else
                @printf("%3.0d %16.8e %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e %9.2e %7.0d %7.0d %8.2f\n", solver.iter, obj, DIMACS_error, err[1], err[2], err[3], err[4], err[5], err[6], solver.cg_iter_pre, solver.cg_iter_cor, solver.itertime)
            end
        elseif solver.kit == 0
            @printf("%3.0d %16.8e %9.2e %8.2f\n", solver.iter, obj, DIMACS_error, solver.itertime)
        else
            @printf("%3.0d %16.8e %9.2e %9.0d %8.2f\n", solver.iter, obj, DIMACS_error, solver.cg_iter_pre + solver.cg_iter_cor, solver.itertime)
        end
    end
    if DIMACS_error < solver.eDIMACS
        solver.status = 1
        if solver.verb > 0
            println("Primal objective: ", obj)
            println("Dual objective:   ", -trCX[] - dx[])
        end
    end
    if DIMACS_error > 1e55
        solver.status = 2
        solver.verb > 0 && println("WARNING: Problem probably infeasible (stopping status = 2)")
    elseif abs(by[]) > 1e55
        solver.status = 3
        solver.verb > 0 && println("WARNING: Problem probably unbounded or infeasible (stopping status = 3)")
    end
end

function S.solve(solver::S.MySolver{Float64}, halpha::S.Halpha)            # src/Solvers.jl:304-361
    ENABLED[] || return invoke(S.solve, Tuple{S.MySolver,S.Halpha}, solver, halpha)     # not enabled: the reference's CPU path
    invoke(S.solve, Tuple{S.MySolver,S.Halpha}, solver, halpha)            # the reference loop; every call inside dispatches to the methods above
    download_solution!(solver)                                             # host arrays for the MOI getters
    return nothing
end

"Float64 is the only GPU element type: every other `MySolver{T}` is rejected instead of silently running on the CPU."
function S.solve(solver::S.MySolver{T}, halpha::S.Halpha) where {T<:AbstractFloat}
    ENABLED[] || return invoke(S.solve, Tuple{S.MySolver,S.Halpha}, solver, halpha)
    throw(ArgumentError("Loraine.Optimizer{$T}: the B200 path supports Float64 only (no Float64xN fallback); " *
                        "do not call LoraineB200.enable! to keep the reference's CPU path for this element type"))
end

"Device phase timers under the reference's TimerOutputs section names (src/makeBBBB.jl:2,30; src/prepare_W.jl:37; src/Solvers.jl:583,676)."
function device_timers(solver)
    ms = zeros(12); calls = zeros(Int64, 12)
    check(dev(solver), ccall((:lrn_timers, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int32), dev(solver).h, ms, calls, 0), "lrn_timers")
    names = ["prepare W", "residuals", solver.datarank == -1 ? "BBBB_rank1" : "BBBBs", "RHS", "cholesky", "solve", "find_step",
             "prec", "CG (Ax + prec apply)", "check_convergence", "prep W SVD", "eigmin"]
    return Dict(names[i] => (ms = ms[i], calls = calls[i]) for i in 1:12)
end

end # module
