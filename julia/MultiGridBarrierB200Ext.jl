# MultiGridBarrierB200Ext.jl -- the reference-side binding of libmgbx.so (include/mgbx.h).
#
# STATUS: an UNEXECUTED binding.  It is written against MultiGridBarrier.jl v1.1.0 exactly as it lies under
# /root/reference -- it reads only fields and calls only functions that exist there (cited below, file:line) -- but the
# build image of this repository has no Julia, so it has never run.  The same C entry points are exercised, call for
# call, by the ctypes binding (multigridbarrier.jl_b200/native.py + solver.py) and its tests.  See INTEGRATION.md.
#
#   using MultiGridBarrier
#   include("MultiGridBarrierB200Ext.jl"); using .MultiGridBarrierB200Ext
#   prob = assemble(amg(subdivide(fem2d_P1(), 10)); p = 1.5)
#   sol  = mgb_solve(prob; device = B200Device)              # same MGBSOL as mgb_solve(prob)
#
# How it plugs in (no change to the reference):
#   mgb_solve(prob; device)                       src/mgb.jl:798-842
#     prob = native_to_device(device, prob)       src/mgb.jl:805, src/device.jl:33-40   <- method added here for B200Device
#     mgb_driver(prob.M, prob.f, prob.g, prob.Q)  src/mgb.jl:831                         <- method added here for M::MgbxPair
#     device_to_native(device, sol)               src/mgb.jl:841, src/device.jl:42-49   <- method added here (identity: z is host)
# What the backend is handed is the reference's own pure-data problem (src/mgb.jl:666-674): the AMG pair with
# `R_fine`, `D_fine`, `w` (src/multigrid.jl:278-288) -- NO level-to-level transfers (they are discarded at
# src/multigrid.jl:166-170) -- the grids f, g and the Convex with its functors and `args` (src/convex.jl:80-86).
# libmgbx recovers the transfers and the per-variable column offsets from R_fine itself (mgbx_amg.T = var_offsets = NULL).
module MultiGridBarrierB200Ext

using MultiGridBarrier, SparseArrays, LinearAlgebra
const MGB = MultiGridBarrier
const LIB = get(ENV, "LIBMGBX", "libmgbx.so")

export B200Device

"Device marker (src/device.jl:18): `mgb_solve(prob; device = B200Device)`."
struct B200Device <: MGB.Device end

# ---------------------------------------------------------------------------------------------- C structs (mgbx.h)
struct Csr;   rows::Int64; cols::Int64; rowptr::Ptr{Int64}; colind::Ptr{Int64}; val::Ptr{Float64}; end
struct Piece; kind::Int32; ni::Int32; nc::Int32; idx::Ptr{Int32}; A::Ptr{Float64}; b::Ptr{Float64};
              p::Ptr{Float64}; mu::Ptr{Float64}; end
struct CConvex; npieces::Int32; pieces::Ptr{Piece}; select::Ptr{Float64}; end
struct CAmg;  n::Int64; N::Int64; p::Int32; nu::Int32; nD::Int32; L::Int32; w::Ptr{Float64}; nops::Int32;
              op_data::Ptr{Ptr{Float64}}; D_var::Ptr{Int32}; D_op::Ptr{Int32}; R_fine::Ptr{Csr}; T::Ptr{Csr};
              var_offsets::Ptr{Int64}; n_global::Int64; var_local::Ptr{Int32}; end
struct CProblem; amg1::CAmg; amg2::CAmg; f_grid::Ptr{Float64}; g_grid::Ptr{Float64}; Q::CConvex;
              barrier_weights::Ptr{Float64}; end
Base.@kwdef mutable struct StepOpts            # mgbx_step_opts; defaults = mgbx_default_step_opts = src/mgb.jl:360-363
    maxit::Int32 = 10000; max_newton::Int32 = 8; initial_step::Int32 = 0; stop_kind::Int32 = 1
    stop_lambda_tol::Float64 = 0.0; stop_theta::Float64 = 0.9; finalize::Int32 = 0; finalize_theta::Float64 = 0.9
    line_search::Int32 = 0; ls_beta::Float64 = 0.5; ls_c1::Float64 = 0.1
end
mutable struct StepResult                      # mgbx_step_result
    converged::Int32; its::NTuple{32,Int32}; y::Float64; gnorm::Float64; inc::Float64
    f01_evals::Int32; f2_evals::Int32; linear_solves::Int32; pcg_iters::Int32
    ms_f01::Float64; ms_f2::Float64; ms_solve::Float64; solve_failures::Int32; its_finalize::Int32; direct_fallbacks::Int32
    StepResult() = new(0, ntuple(_ -> Int32(0), 32), 0.0, 0.0, 0.0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0, 0, 0)
end
mutable struct ScalarsOut                      # mgbx_scalars_out
    c_dot_Dz::Float64; var_max::NTuple{12,Float64}; var_absmax::NTuple{12,Float64}; all_finite::Int32
    ScalarsOut() = new(0.0, ntuple(_ -> 0.0, 12), ntuple(_ -> 0.0, 12), 0)
end
const ZERO_AMG = CAmg(0, 0, 0, 0, 0, 0, C_NULL, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0, C_NULL)

lasterr(h) = unsafe_string(ccall((:mgbx_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(rc, h) = rc < 0 ? error("libmgbx error $rc: " * lasterr(h)) : rc

# ---------------------------------------------------------------------------------------------- packing
"CSR arrays (0-based, Int64) of a CPU matrix: the CSR of R is the CSC of R' (SparseMatrixCSC fields colptr/rowval/nzval)."
function csr!(keep, A::AbstractMatrix)
    At = sparse(transpose(sparse(A)))                       # dense spectral R_fine go through sparse() too
    rp = Int64.(At.colptr .- 1); ci = Int64.(At.rowval .- 1); v = Float64.(At.nzval)
    push!(keep, rp, ci, v)
    Csr(size(A, 1), size(A, 2), pointer(rp), pointer(ci), pointer(v))
end

"""One operator row of `AMG.D_fine` (src/multigrid.jl:505-510) -> (state variable, p x p x N block array).
FEM: `BlockColumn` (src/BlockMatrices.jl:38-44: fields active_block::BlockDiag (p, q, N, data), active_col, nu).
Spectral: a dense n x (nu n) matrix with one non-zero n x n block (`_block_column`, src/BlockMatrices.jl:666-672): N = 1, p = n."""
function operator_block(Dk, n::Int)
    if Dk isa MGB.BlockColumn
        blk = Dk.active_block
        return Dk.active_col, Dk.nu, blk.p, blk.N, blk.data
    end
    nu = size(Dk, 2) ÷ n
    a = something(findfirst(j -> any(!iszero, view(Dk, :, (j-1)*n+1:j*n)), 1:nu), 1)
    return a, nu, n, 1, reshape(Matrix{Float64}(Dk[:, (a-1)*n+1:a*n]), n, n, 1)
end

is_identity_blocks(data) = all(e -> view(data, :, :, e) == I, axes(data, 3))

"`AMG` (src/multigrid.jl:278-288) -> mgbx_amg.  Only R_fine, D_fine and w are read: T and var_offsets stay NULL."
function pack_amg!(keep, M)
    n = length(M.w)
    datas = Any[]; ptrs = Ptr{Float64}[]; D_var = Int32[]; D_op = Int32[]
    nu = 0; p = 0; N = 0
    for Dk in M.D_fine
        a, nu, p, N, data = operator_block(Dk, n)
        push!(D_var, a - 1)
        if is_identity_blocks(data)
            push!(D_op, -1)                                   # :id is a flag for the library, never uploaded
        else
            k = findfirst(d -> d === data || d == data, datas)
            if k === nothing
                d64 = Array{Float64,3}(data)
                push!(datas, d64); push!(ptrs, pointer(d64)); k = length(datas)
            end
            push!(D_op, k - 1)
        end
    end
    L = length(M.R_fine)
    Rs = [csr!(keep, M.R_fine[l]) for l in 1:L]
    w = Vector{Float64}(M.w)
    push!(keep, datas, ptrs, D_var, D_op, Rs, w)
    CAmg(n, N, p, nu, length(M.D_fine), L, pointer(w), length(ptrs), isempty(ptrs) ? C_NULL : pointer(ptrs), pointer(D_var), pointer(D_op),
         pointer(Rs), C_NULL, C_NULL, 0, C_NULL)
end

# idx of a convex functor: the EP functors carry it as a field (src/convex_euclidian_power.jl:71-73), the convex_linear
# closures capture the variable `idx` (src/convex_linear.jl:119-214), which Julia exposes as a field of the closure.
functor_idx(F) = getfield(F, :idx)
idx_vector(idx::Colon, ni) = nothing                          # Colon = the leading ni rows (mgbx_piece.idx = NULL)
idx_vector(idx, ni) = Int32.(collect(idx) .- 1)

"Convex (src/convex.jl:80-86) -> mgbx_convex: pieces = EP / LINEAR descriptors, grids = Q.args (n x k matrices, column-major)."
function pack_convex!(keep, Q)
    pieces = Piece[]
    grid(a) = (g = Matrix{Float64}(reshape(a, size(a, 1), :)); push!(keep, g); g)
    function add_piece(F, args)
        if F isa MGB.EuclidianPowerBarrier                     # args = (A_grid n x nz^2, b_grid n x nz, p_grid, mu_grid), :446-452
            A, b, p, mu = grid(args[1]), grid(args[2]), grid(args[3]), grid(args[4])
            nz = size(b, 2)
            iv = idx_vector(functor_idx(F), nz); iv === nothing || push!(keep, iv)
            push!(pieces, Piece(0, nz, nz, iv === nothing ? C_NULL : pointer(iv), pointer(A), pointer(b), pointer(p), pointer(mu)))
            return 4
        else                                                   # convex_linear: args = (A_grid n x (nc ni), b_grid n x nc), :216-222
            A, b = grid(args[1]), grid(args[2])
            nc = size(b, 2); ni = size(A, 2) ÷ nc
            iv = idx_vector(functor_idx(F), ni); iv === nothing || push!(keep, iv)
            push!(pieces, Piece(1, ni, nc, iv === nothing ? C_NULL : pointer(iv), pointer(A), pointer(b), C_NULL, C_NULL))
            return 2
        end
    end
    F0 = Q.barrier[1]
    sel = Ptr{Float64}(C_NULL)
    if F0 isa MGB.PiecewiseBarrierF0                           # args = (select n x K, piece_1 args..., piece_K args...), src/convex_piecewise.jl:158-160
        s = grid(Q.args[1]); sel = pointer(s)
        k = 2
        for Fk in F0.barrier_f0s
            k += add_piece(Fk, Q.args[k:end])
        end
    else
        add_piece(F0, Q.args)
    end
    push!(keep, pieces)
    CConvex(length(pieces), pointer(pieces), sel)
end

# ---------------------------------------------------------------------------------------------- native_to_device
"What `native_to_device(B200Device, prob)` puts into `prob.M`: the CPU AMG pair, wrapped so that `mgb_driver` dispatches here."
struct MgbxPair{MT}
    M::MT
end
Base.getindex(P::MgbxPair, k) = P.M[k]
Base.length(P::MgbxPair) = length(P.M)

function MGB.native_to_device(::Type{B200Device}, prob::MGB.MGBProblem{T}) where {T}
    T === Float64 || error("B200Device: libmgbx computes in Float64 (got $T)")
    MGB.MGBProblem{T}(MgbxPair(prob.M), prob.f, prob.g, prob.Q, prob.geometry)
end
MGB.device_to_native(::Type{B200Device}, sol) = sol            # z and the SOL tuples are host arrays already

mutable struct Handle
    h::Ptr{Cvoid}; keep::Vector{Any}; n::Int; nu::Int; L::NTuple{2,Int}; feas::CAmg
end

function Handle(M, f, g, Q, bw)
    keep = Any[]; n, nu = size(g)
    a1 = pack_amg!(keep, M[1]); a2 = pack_amg!(keep, M[2])
    q = pack_convex!(keep, Q)
    fg = Matrix{Float64}(f); gg = Matrix{Float64}(g); push!(keep, fg, gg)
    bwp = bw === nothing ? Ptr{Float64}(C_NULL) : (b = Vector{Float64}(bw); push!(keep, b); pointer(b))
    cp = Ref(CProblem(a1, ZERO_AMG, pointer(fg), pointer(gg), q, bwp))     # the feasibility AMG is attached only if phase I runs
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = GC.@preserve keep ccall((:mgbx_create, LIB), Cint, (Ptr{CProblem}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), cp, C_NULL, out)
    rc == 0 || error("mgbx_create: " * lasterr(C_NULL))
    Handle(out[], keep, n, nu, (length(M[1].R_fine), length(M[2].R_fine)), a2)
end
close!(H::Handle) = (H.h == C_NULL || ccall((:mgbx_destroy, LIB), Cvoid, (Ptr{Cvoid},), H.h); H.h = C_NULL)

function step!(H::Handle, which, t, o::StepOpts)
    r = StepResult()
    rc = check(ccall((:mgbx_step, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ref{StepOpts}, Ref{StepResult}), H.h, which, t, o, r), H.h)
    rc == 2 && error("newton: non-finite objective, gradient or direction at t=$t")
    (converged = rc == 0, its = Int.(collect(r.its)[1:H.L[which+1]]), its_finalize = Int(r.its_finalize))
end
scalars(H::Handle, which) = (s = ScalarsOut(); check(ccall((:mgbx_scalars, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{ScalarsOut}), H.h, which, s), H.h); s)
function get_z(H::Handle, which, nu; unfinalized = false)
    z = Matrix{Float64}(undef, H.n, nu)
    sym = unfinalized ? :mgbx_get_z_unfinalized : :mgbx_get_z
    check(unfinalized ? ccall((:mgbx_get_z_unfinalized, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), H.h, which, z) :
                        ccall((:mgbx_get_z, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), H.h, which, z), H.h)
    z
end

# ---------------------------------------------------------------------------------------------- solver options
# The reference passes closures; their captured variables are fields of the closure objects.
function stop_options(sc)                                      # stopping_inexact / stopping_exact, src/newton.jl:187,222-225
    hasproperty(sc, :lambda_tol) && return (1, Float64(sc.lambda_tol), Float64(sc.exact_stop.theta))
    hasproperty(sc, :theta) && return (0, 0.0, Float64(sc.theta))
    error("B200Device: stopping_criterion must come from stopping_inexact or stopping_exact")
end
function linesearch_options(ls)                                # linesearch_backtracking / linesearch_illinois, src/newton.jl:84-103,139-154
    hasproperty(ls, :c1) && return (0, Float64(ls.beta), Float64(ls.c1))
    hasproperty(ls, :beta) && return (1, Float64(ls.beta), 0.1)
    error("B200Device: line_search must come from linesearch_backtracking or linesearch_illinois")
end

"src/mgb.jl:91-183 over a handle: mgb_step -> mgbx_step, dot(w .* c, Dz) -> mgbx_scalars, early_stop(z, t) -> early_stop(t)."
function core(H::Handle, which; tol = sqrt(eps()), t = 0.1, maxit = 10000, kappa = 10.0, early_stop = t -> false,
              max_newton = Int(ceil(log2(-log2(eps())))) + 2, finalize_theta = nothing, stop, ls, progress = x -> nothing,
              printlog = (x...) -> nothing)
    target = 1 / tol; kappa0 = kappa; t_begin = time()
    opts(tt, initial) = StepOpts(maxit = maxit, max_newton = max_newton, initial_step = initial, stop_kind = stop[1], stop_lambda_tol = stop[2],
                                 stop_theta = stop[3], finalize = (finalize_theta !== nothing && tt >= target) ? 1 : 0,
                                 finalize_theta = something(finalize_theta, 0.9), line_search = ls[1], ls_beta = ls[2], ls_c1 = ls[3])
    S = step!(H, which, t, opts(t, 1))
    S.converged || throw(MGB.MGBConvergenceFailure("Initial centering failed in mgb_solve at t=$t, tol=$tol, maxit=$maxit.", :stall))
    its = [S.its]; ts = [t]; kappas = [kappa]; cdz = [scalars(H, which).c_dot_Dz]; times = [time()]; k = 1
    while t < target && kappa > 1 && k < maxit && !early_stop(t)
        k += 1; itk = zeros(Int, length(S.its))
        while kappa > 1
            t1 = kappa * t; S = step!(H, which, t1, opts(t1, 0)); itk .+= S.its
            if S.converged
                maximum(S.its) <= max_newton / 2 && (kappa = min(kappa0, kappa^2))
                t = t1; break
            end
            printlog("mgb_core: t refinement failed, shrinking kappa")
            kappa = sqrt(kappa)
        end
        push!(its, itk); push!(ts, t); push!(kappas, kappa); push!(cdz, scalars(H, which).c_dot_Dz); push!(times, time())
        progress(min(1.0, log(t / ts[1]) / log(target / ts[1])))
    end
    (t >= target || early_stop(t)) || throw(MGB.MGBConvergenceFailure(
        "Convergence failure in mgb_solve at t=$t, k=$k, kappa=$kappa, tol=$tol, maxit=$maxit.", kappa <= 1 ? :stall : :iteration_limit))
    t_end = time()
    nuw = which == 0 ? H.nu : H.nu + 1
    z = vec(get_z(H, which, nuw)); zu = vec(get_z(H, which, nuw; unfinalized = true))
    (z = z, z_unfinalized = zu, its = reduce(hcat, its), ts = ts, kappas = kappas, t_begin = t_begin, t_end = t_end,
     t_elapsed = t_end - t_begin, times = times, c_dot_Dz = cdz)
end

"""`mgb_driver` (src/mgb.jl:332-584) for a problem moved to `B200Device`: feasibility probe, phase I with box escalation,
handoff, `_matched_t`, main ramp -- control flow and scalars here, every array operation inside libmgbx."""
function MGB.mgb_driver(P::MgbxPair, f, g, Q::MGB.Convex{T};
        t = T(0.1), t_feasibility = t, feasibility_Rmax = one(T) / sqrt(eps(T)), progress = x -> nothing,
        stopping_criterion = MGB.stopping_inexact(T(0.25) / sqrt(T(length(P[1].w))), T(0.9)),
        printlog = (args...) -> nothing, line_search = MGB.linesearch_backtracking(T),
        finalize = MGB.stopping_exact(T(0.9)), barrier_nodes = .!iszero.(P[1].w), rest...) where {T}
    M = P.M
    bw = MGB._barrier_weights(M[1].w, barrier_nodes)           # src/convex.jl:279-304 (nothing = plain 1/n average)
    stop = stop_options(stopping_criterion); ls = linesearch_options(line_search)
    fth = (finalize === false || finalize isa MGB.NoFinalize) ? nothing : stop_options(finalize)[3]
    H = Handle(M, f, g, Q, bw)
    try
        SOL_feas = nothing
        need = Ref{Int32}(0); b = Ref(0.0); zabs = Ref(0.0)
        p1() = check(ccall((:mgbx_phase1_init, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Float64}, Ref{Float64}), H.h, need, b, zabs), H.h)
        p1()
        if need[] != 0
            GC.@preserve H check(ccall((:mgbx_attach_feasibility, LIB), Cint, (Ptr{Cvoid}, Ref{CAmg}), H.h, H.feas), H.h)
            p1()
            Rbox = max(10.0, 10zabs[]); Rmax = max(Float64(feasibility_Rmax), Rbox); first = true
            feasible() = scalars(H, 1).var_max[H.nu+1] < 0
            while true
                printlog("mgb_driver: feasibility phase with bounding box R=", Rbox)
                check(ccall((:mgbx_set_feasibility_box, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), H.h, b[], Rbox), H.h)
                first || check(ccall((:mgbx_reset_feasibility_state, LIB), Cint, (Ptr{Cvoid},), H.h), H.h)   # no warm start, :538-543
                first = false
                tfirst = Inf
                stopfn(tt) = feasible() ? (tfirst = min(tfirst, tt); tt >= 2tfirst) : false                   # handoff rule, :486-491
                failed = false
                try
                    SOL_feas = core(H, 1; t = Float64(t_feasibility), early_stop = stopfn, stop = stop, ls = ls, finalize_theta = fth,
                                    printlog = printlog, rest...)
                catch e
                    e isa InterruptException && rethrow()
                    printlog("mgb_driver: feasibility solve failed at R=", Rbox, ": ", e)
                    failed = true
                end
                if !failed
                    feasible() && break
                    sc = scalars(H, 1); vmax = maximum(sc.var_absmax[1:H.nu]); smax = sc.var_max[H.nu+1]
                    vmax <= Rbox / 2 && throw(MGB.MGBConvergenceFailure(
                        "The problem appears to be infeasible: the feasibility subproblem converged to a minimizer with positive " *
                        "constraint violation (max slack ~ $smax) strictly inside the bounding box (max |nodal value| ~ $vmax <= R/2 with R = $Rbox).",
                        :infeasible))
                    printlog("mgb_driver: phase-I minimizer presses the box; growing R")
                end
                10Rbox > Rmax && throw(MGB.MGBConvergenceFailure(
                    "Could not find a strictly feasible point with nodal values bounded by R = $Rbox (cap feasibility_Rmax ~ $Rmax).",
                    :feasibility_Rmax))
                Rbox *= 10
            end
            check(ccall((:mgbx_handoff, LIB), Cint, (Ptr{Cvoid},), H.h), H.h)
            tm = Ref(0.0); tstar = Ref(0.0)
            check(ccall((:mgbx_matched_t, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ref{Float64}, Ref{Float64}), H.h, Float64(t), tm, tstar), H.h)
            t = min(t, T(tm[]))
        end
        SOL_main = core(H, 0; t = Float64(t), stop = stop, ls = ls, finalize_theta = fth, progress = progress, printlog = printlog, rest...)
        z = get_z(H, 0, H.nu)
        return (; z, SOL_feasibility = SOL_feas, SOL_main)
    finally
        close!(H)
    end
end

end # module
