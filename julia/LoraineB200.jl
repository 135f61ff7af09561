# LoraineB200.jl -- `ccall` shims that put libloraine_b200.so behind the unchanged Loraine.Optimizer / MOI / JuMP surface.
#
# Usage (on a machine with Julia, Loraine.jl v0.2.5 and a B200):
#     using Loraine; include("LoraineB200.jl"); LoraineB200.enable!("/path/to/libloraine_b200.so")
# After `enable!`, `Loraine.Solvers.solve` for `MySolver{Float64}` runs the reference's own control flow
# (src/Solvers.jl:304-361, :448-478; src/predictor_corrector.jl) with every array expression replaced by one ccall.
# `Optimizer{T}` with T != Float64 throws an ArgumentError (no fallback).
#
# NOTE: Julia is not available in the build container; this file was written against the reference sources and the C
# header and has not been executed.  loraine.jl_b200/solver.py is the executed mirror of the same logic.
module LoraineB200

using Loraine
using SparseArrays, LinearAlgebra, Printf
const S = Loraine.Solvers

const LIB = Ref{String}("libloraine_b200.so")

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

"Create the device handle and upload the prepared model (outputs of `_prepare_A`, src/model.jl:120-150)."
function attach!(solver::S.MySolver{Float64})
    md = solver.model
    opt = Ref(Options(solver.kit, solver.datarank, solver.preconditioner, solver.erank, solver.aamat, solver.datasparsity,
                      0, 0, 0.0, 0.0, -1, 0))
    hr = Ref{Ptr{Cvoid}}(C_NULL)
    ms = Int64.(md.msizes)
    rc = ccall((:lrn_create, LIB[]), Int32, (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Int64}, Int64, Ref{Options}),
               hr, md.n, md.nlmi, ms, md.nlin, opt)
    rc == 0 || error("loraine_b200: lrn_create failed ($rc); there is no CPU fallback")
    d = Device(hr[])
    for i in 1:md.nlmi
        AAi = md.AA[i]                        # n x m^2 (row k = vec(calA_k))
        GC.@preserve AAi check(d, ccall((:lrn_set_block_AA, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                        d.h, i - 1, csc(AAi)...), "lrn_set_block_AA")
        Ci = SparseMatrixCSC{Float64,Int64}(md.C[i])
        GC.@preserve Ci check(d, ccall((:lrn_set_block_C, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                       d.h, i - 1, csc(Ci)...), "lrn_set_block_C")
        if solver.datarank == -1 && !isempty(md.B)
            Bi = md.B[i]
            GC.@preserve Bi check(d, ccall((:lrn_set_block_B, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                           d.h, i - 1, csc(Bi)...), "lrn_set_block_B")
        end
    end
    if md.nlin > 0
        Cl = SparseMatrixCSC{Float64,Int64}(md.C_lin); dl = Vector{Float64}(md.d_lin)
        GC.@preserve Cl dl check(d, ccall((:lrn_set_lin, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
                                          d.h, csc(Cl)..., dl), "lrn_set_lin")
    end
    b = Vector{Float64}(md.b)
    check(d, ccall((:lrn_set_b, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), d.h, b), "lrn_set_b")
    check(d, ccall((:lrn_finalize, LIB[]), Int32, (Ptr{Cvoid},), d.h), "lrn_finalize")
    DEVICES[solver] = d
    return d
end

dev(solver) = DEVICES[solver]

function upload_iterate!(solver)          # after initial_point (src/initial_point.jl)
    d = dev(solver); md = solver.model
    X = [Matrix{Float64}(x) for x in solver.X]; Sm = [Matrix{Float64}(x) for x in solver.S]
    Xp = [pointer(x) for x in X]; Sp = [pointer(x) for x in Sm]
    y = vec(Float64.(solver.y)); xl = vec(Float64.(solver.X_lin)); sl = vec(Float64.(solver.S_lin))
    GC.@preserve X Sm y xl sl check(d, ccall((:lrn_set_iterate, LIB[]), Int32,
        (Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        d.h, Xp, Sp, y, md.nlin > 0 ? pointer(xl) : C_NULL, md.nlin > 0 ? pointer(sl) : C_NULL), "lrn_set_iterate")
end

function download_solution!(solver)       # MOI getters read solver.y / X / X_lin (src/MOI_wrapper.jl:315-354)
    d = dev(solver); md = solver.model
    y = zeros(md.n); X = [zeros(Int(m), Int(m)) for m in md.msizes]; xl = zeros(md.nlin)
    Xp = [pointer(x) for x in X]
    GC.@preserve X y xl check(d, ccall((:lrn_get_solution, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Ptr{Float64}}, Ptr{Float64}),
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

call0(solver, f, what) = check(dev(solver), ccall((f, LIB[]), Int32, (Ptr{Cvoid},), dev(solver).h), what)

# --- hot-path replacements: same names / argument meaning as module Solvers ---------------------------------------------
function find_mu(solver)
    mu = Ref(0.0)
    check(dev(solver), ccall((:lrn_find_mu, LIB[]), Int32, (Ptr{Cvoid}, Ref{Float64}), dev(solver).h, mu), "lrn_find_mu")
    solver.mu = mu[]
end

function prepare_W(solver)
    st4 = Ref{Int32}(0)
    check(dev(solver), ccall((:lrn_prepare_W, LIB[]), Int32, (Ptr{Cvoid}, Ref{Int32}), dev(solver).h, st4), "lrn_prepare_W")
    st4[] != 0 && (solver.status = 4)
end

function pcg(solver, kind)
    it = Ref{Int64}(0); code = Ref{Int32}(0)
    check(dev(solver), ccall((:lrn_pcg, LIB[]), Int32, (Ptr{Cvoid}, Float64, Int64, Int32, Ref{Int64}, Ref{Int32}),
                             dev(solver).h, Float64(solver.tol_cg), 10000, kind, it, code), "lrn_pcg")
    return it[]
end

function find_step(solver)
    d = dev(solver); md = solver.model
    a = zeros(max(1, md.nlmi)); b = zeros(max(1, md.nlmi)); al = Ref(1.0); bl = Ref(1.0)
    check(d, ccall((:lrn_find_step, LIB[]), Int32,
                   (Ptr{Cvoid}, Int32, Float64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Float64}),
                   d.h, solver.predict ? 1 : 0, Float64(solver.sigma), Float64(solver.mu), Float64(solver.tau), a, b, al, bl),
          "lrn_find_step")
    solver.alpha = a[1:md.nlmi]; solver.beta = b[1:md.nlmi]; solver.alpha_lin = al[]; solver.beta_lin = bl[]
end

function predictor(solver, halpha)                       # src/predictor_corrector.jl:5-146, control flow kept
    solver.predict = true
    call0(solver, :lrn_residuals, "lrn_residuals")                         # :8-22
    solver.kit == 0 && call0(solver, :lrn_schur_assemble, "lrn_schur_assemble")   # :24-40
    call0(solver, :lrn_rhs_predictor, "lrn_rhs_predictor")                 # :43-50
    if solver.kit == 0
        try
            call0(solver, :lrn_schur_factor, "lrn_schur_factor")           # :57
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
        kind != 0 && check(dev(solver), ccall((:lrn_prec_prepare, LIB[]), Int32, (Ptr{Cvoid}, Int32), dev(solver).h, kind), "lrn_prec_prepare")
        n = pcg(solver, kind)
        solver.cg_iter_pre += n; solver.cg_iter_tot += n
    end
    find_step(solver)
end

function sigma_update(solver)                            # src/predictor_corrector.jl:148-179
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

function corrector(solver, halpha)                       # src/predictor_corrector.jl:181-246
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
    find_step(solver)
end

function check_convergence_terms!(solver)                # arithmetic of src/Solvers.jl:496-523; the printing / status logic stays
    err = zeros(6); by = Ref(0.0); trCX = Ref(0.0); dx = Ref(0.0)
    check(dev(solver), ccall((:lrn_dimacs, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ref{Float64}, Ref{Float64}),
                             dev(solver).h, err, by, trCX, dx), "lrn_dimacs")
    solver.err1, solver.err2, solver.err3, solver.err4, solver.err5, solver.err6 = err
    return by[], trCX[], dx[]
end

"Route Loraine.Solvers' hot-path methods for `MySolver{Float64}` to the device and reject other element types."
function enable!(libpath::AbstractString = LIB[])
    LIB[] = libpath
    @eval Loraine begin
        function Optimizer{T}() where {T}       # src/MOI_wrapper.jl:52-66 -- Float64 only
            T === Float64 || throw(ArgumentError("Loraine.Optimizer{$T}: the B200 path supports Float64 only (no Float64xN fallback)"))
            return invoke_original_optimizer(T)
        end
    end
    @eval Loraine.Solvers begin
        find_mu(s::MySolver{Float64}) = Main.LoraineB200.find_mu(s)
        prepare_W(s::MySolver{Float64}) = Main.LoraineB200.prepare_W(s)
        predictor(s::MySolver{Float64}, ha::Halpha) = Main.LoraineB200.predictor(s, ha)
        sigma_update(s::MySolver{Float64}) = Main.LoraineB200.sigma_update(s)
        corrector(s::MySolver{Float64}, ha) = Main.LoraineB200.corrector(s, ha)
        find_step(s::MySolver{Float64}) = Main.LoraineB200.find_step(s)
    end
    return nothing
end

end # module
