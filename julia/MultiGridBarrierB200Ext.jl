# MultiGridBarrierB200Ext.jl -- the reference-side binding of libmgbx.so (include/mgbx.h).
#
# STATUS: written against MultiGridBarrier.jl v1.1.0 (the reference under /root/reference) but NOT executed in this
# repository's build image (no Julia toolchain there); the same entry points are exercised by the ctypes binding
# multigridbarrier.jl_b200/native.py + solver.py, which this file mirrors function by function.  See INTEGRATION.md.
#
#   using MultiGridBarrier, MultiGridBarrierB200Ext
#   mg   = amg(subdivide(fem2d_P1(), 10))
#   prob = assemble(mg; p = 1.5)
#   sol  = mgb_solve_b200(mg, prob)                    # same MGBSOL as mgb_solve(prob)
module MultiGridBarrierB200Ext

using MultiGridBarrier, SparseArrays, LinearAlgebra
const MGB = MultiGridBarrier
const LIB = get(ENV, "LIBMGBX", "libmgbx.so")

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
Base.@kwdef mutable struct StepOpts
    maxit::Int32 = 10000; max_newton::Int32 = 8; initial_step::Int32 = 0; stop_kind::Int32 = 1
    stop_lambda_tol::Float64 = 0.0; stop_theta::Float64 = 0.9; finalize::Int32 = 0; finalize_theta::Float64 = 0.9
    line_search::Int32 = 0; ls_beta::Float64 = 0.5; ls_c1::Float64 = 0.1
end
mutable struct StepResult
    converged::Int32; its::NTuple{32,Int32}; y::Float64; gnorm::Float64; inc::Float64
    f01_evals::Int32; f2_evals::Int32; linear_solves::Int32; pcg_iters::Int32
    ms_f01::Float64; ms_f2::Float64; ms_solve::Float64
    StepResult() = new(0, ntuple(_ -> Int32(0), 32), 0.0, 0.0, 0.0, 0, 0, 0, 0, 0.0, 0.0, 0.0)
end
mutable struct ScalarsOut
    c_dot_Dz::Float64; var_max::NTuple{12,Float64}; var_absmax::NTuple{12,Float64}; all_finite::Int32
    ScalarsOut() = new(0.0, ntuple(_ -> 0.0, 12), ntuple(_ -> 0.0, 12), 0)
end

lasterr(h) = unsafe_string(ccall((:mgbx_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(rc, h) = rc < 0 ? error("libmgbx error $rc: " * lasterr(h)) : rc

# ---------------------------------------------------------------------------------------------- packing
"CSC -> CSR (0-based, int64): the CSR arrays of A are the CSC arrays of A'."
function csr!(keep, A::AbstractMatrix)
    At = sparse(transpose(sparse(A)))
    rp = Int64.(At.colptr .- 1); ci = Int64.(At.rowval .- 1); v = Float64.(At.nzval)
    push!(keep, rp, ci, v)
    Csr(size(A, 1), size(A, 2), pointer(rp), pointer(ci), pointer(v))
end
dimsonly(A) = Csr(size(A, 1), size(A, 2), C_NULL, C_NULL, C_NULL)

"Level transfers T[l] with sub[l+1] * T[l] = refine[l] * sub[l] (nested spaces: exact), per state variable, joined block-diagonally."
function transfers(mg, state_variables)
    L = length(first(values(mg.refine)))
    Ts = Vector{SparseMatrixCSC{Float64,Int}}(undef, L - 1)
    for l in 1:L-1
        blocks = map(eachrow(state_variables)) do sv
            X = sv[2]; s1 = sparse(mg.subspaces[X][l+1]); rhs = sparse(mg.refine[X][l] * mg.subspaces[X][l])
            G = s1' * s1
            sparse(G \ Matrix(s1' * rhs))            # small per-variable normal equations; diagonal for gluing matrices
        end
        Ts[l] = blockdiag(blocks...)
    end
    Ts
end

function pack_amg!(keep, M::MGB.AMG, T, nu::Int)
    n = length(M.w); D = M.D_fine
    blk = D[1].active_block; p = blk.p; N = blk.N
    names = Ptr{Float64}[]; datas = Any[]; D_var = Int32[]; D_op = Int32[]
    for Dk in D
        data = Dk.active_block.data
        push!(D_var, Dk.active_col - 1)
        if all(e -> data[:, :, e] == I, 1:min(N, 4)) && data == repeat(Matrix{Float64}(I, p, p), 1, 1, N)
            push!(D_op, -1)                                   # :id is a flag, never uploaded
        else
            k = findfirst(d -> d === data, datas)
            k === nothing && (push!(datas, data); push!(names, pointer(data)); k = length(datas))
            push!(D_op, k - 1)
        end
    end
    L = length(M.R_fine)
    Rs = [l < L ? dimsonly(M.R_fine[l]) : csr!(keep, M.R_fine[l]) for l in 1:L]
    Tc = [csr!(keep, T[l]) for l in 1:L-1]
    # var_offsets[l][k]: first column of variable k at level l -- from the per-variable block widths of R_fine[l]
    voff = Int64[]
    for l in 1:L
        widths = MGB.block_col_sizes(M.R_fine[l], nu)        # helper the shim adds next to amg_helper (multigrid.jl:474-512)
        append!(voff, cumsum([0; widths]))
    end
    push!(keep, names, datas, D_var, D_op, Rs, Tc, voff)
    CAmg(n, N, p, nu, length(D), L, pointer(M.w), length(names), pointer(names), pointer(D_var), pointer(D_op),
         pointer(Rs), isempty(Tc) ? C_NULL : pointer(Tc), pointer(voff), 0, C_NULL)
end

"Convex set -> descriptor: the functor types carry nz / idx (convex_euclidian_power.jl:71-76), Q.args carry the grids."
function pack_convex!(keep, Q, n)
    pieces = Piece[]
    function one(kind, idx, nc, ni, A, b, p, mu)
        idx32 = idx === nothing ? C_NULL : (v = Int32.(collect(idx) .- 1); push!(keep, v); pointer(v))
        push!(keep, A, b, p, mu)
        push!(pieces, Piece(kind, ni, nc, idx32, pointer(A), pointer(b), p === nothing ? C_NULL : pointer(p),
                            mu === nothing ? C_NULL : pointer(mu)))
    end
    F0 = Q.barrier[1]
    if F0 isa MGB.EuclidianPowerBarrier                        # args = (A_grid n x nz^2, b_grid n x nz, p_grid, mu_grid)
        A, b, p, mu = Q.args; nz = size(b, 2)
        one(0, MGB.functor_idx(F0), nz, nz, A, b, p, mu)
        sel = C_NULL
    elseif F0 isa MGB.PiecewiseBarrierF0                       # args = (select, piece_1 args..., piece_K args...)
        sel = pointer(Q.args[1]); k = 2
        for Fk in F0.pieces
            if Fk isa MGB.EuclidianPowerBarrier
                A, b, p, mu = Q.args[k:k+3]; k += 4; nz = size(b, 2); one(0, MGB.functor_idx(Fk), nz, nz, A, b, p, mu)
            else
                A, b = Q.args[k:k+1]; k += 2; nc = size(b, 2); one(1, MGB.functor_idx(Fk), nc, size(A, 2) ÷ nc, A, b, nothing, nothing)
            end
        end
    else                                                       # convex_linear closure: args = (A_grid n x (nc*ni), b_grid n x nc)
        A, b = Q.args; nc = size(b, 2)
        one(1, MGB.functor_idx(F0), nc, size(A, 2) ÷ nc, A, b, nothing, nothing)
        sel = C_NULL
    end
    push!(keep, pieces)
    CConvex(length(pieces), pointer(pieces), sel)
end

# ---------------------------------------------------------------------------------------------- handle
mutable struct Handle
    h::Ptr{Cvoid}; keep::Vector{Any}; n::Int; nu::Int; L::NTuple{2,Int}; feas::Union{Nothing,CAmg}
end

function Handle(mg, prob; barrier_weights = nothing)
    keep = Any[]; nu = size(prob.g, 2); n = size(prob.g, 1)
    sv = MGB.state_variables(prob)                              # the (name, space) table assemble() used
    T1 = transfers(mg, sv); T2 = transfers(mg, vcat(sv, [:feasibility_slack :full]))
    a1 = pack_amg!(keep, prob.M[1], T1, nu); a2 = pack_amg!(keep, prob.M[2], T2, nu + 1)
    q = pack_convex!(keep, prob.Q, n)
    bw = barrier_weights === nothing ? C_NULL : (push!(keep, barrier_weights); pointer(barrier_weights))
    zero_amg = CAmg(0, 0, 0, 0, 0, 0, C_NULL, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, 0, C_NULL)
    cp = Ref(CProblem(a1, zero_amg, pointer(prob.f), pointer(prob.g), q, bw))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mgbx_create, LIB), Cint, (Ptr{CProblem}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), cp, C_NULL, out)
    rc == 0 || error("mgbx_create: " * lasterr(C_NULL))
    Handle(out[], keep, n, nu, (length(prob.M[1].R_fine), length(prob.M[2].R_fine)), a2)
end
close!(H::Handle) = (ccall((:mgbx_destroy, LIB), Cvoid, (Ptr{Cvoid},), H.h); H.h = C_NULL)

function step!(H::Handle, which, t, o::StepOpts)
    r = StepResult()
    rc = check(ccall((:mgbx_step, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ref{StepOpts}, Ref{StepResult}), H.h, which, t, o, r), H.h)
    rc == 2 && error("newton: non-finite objective, gradient or direction at t=$t")
    (converged = rc == 0, its = Int.(collect(r.its)[1:H.L[which+1]]))
end
scalars(H::Handle, which) = (s = ScalarsOut(); check(ccall((:mgbx_scalars, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{ScalarsOut}), H.h, which, s), H.h); s)

# ---------------------------------------------------------------------------------------------- mgb_core on a handle
"src/mgb.jl:91-183 with its three data touches replaced: mgb_step -> step!, dot(c, Dz) -> scalars, early_stop(z) -> early_stop(t)."
function core(H::Handle, which; tol = sqrt(eps()), t = 0.1, maxit = 10000, kappa = 10.0, early_stop = t -> false,
              max_newton = Int(ceil(log2(-log2(eps())))) + 2, finalize = true, finalize_theta = 0.9, stop_lambda_tol, line_search = 0)
    target = 1 / tol; kappa0 = kappa
    opts(tt, initial) = StepOpts(maxit = maxit, max_newton = max_newton, initial_step = initial, stop_lambda_tol = stop_lambda_tol,
                                 finalize = (finalize && tt >= target) ? 1 : 0, finalize_theta = finalize_theta, line_search = line_search)
    S = step!(H, which, t, opts(t, 1))
    S.converged || throw(MGB.MGBConvergenceFailure("Initial centering failed in mgb_solve at t=$t, tol=$tol, maxit=$maxit.", :stall))
    its = [S.its]; ts = [t]; kappas = [kappa]; cdz = [scalars(H, which).c_dot_Dz]; k = 1
    while t < target && kappa > 1 && k < maxit && !early_stop(t)
        k += 1; itk = zeros(Int, length(S.its))
        while kappa > 1
            t1 = kappa * t; S = step!(H, which, t1, opts(t1, 0)); itk .+= S.its
            if S.converged
                maximum(S.its) <= max_newton / 2 && (kappa = min(kappa0, kappa^2))
                t = t1; break
            end
            kappa = sqrt(kappa)
        end
        push!(its, itk); push!(ts, t); push!(kappas, kappa); push!(cdz, scalars(H, which).c_dot_Dz)
    end
    (t >= target || early_stop(t)) || throw(MGB.MGBConvergenceFailure(
        "Convergence failure in mgb_solve at t=$t, k=$k, kappa=$kappa, tol=$tol, maxit=$maxit.", kappa <= 1 ? :stall : :iteration_limit))
    (its = reduce(hcat, its), ts = ts, kappas = kappas, c_dot_Dz = cdz)
end

"src/mgb.jl:332-584: feasibility probe, phase I with box escalation, handoff, _matched_t, main ramp."
function driver(H::Handle; t = 0.1, t_feasibility = t, feasibility_Rmax = 1 / sqrt(eps()), printlog = (x...) -> nothing, rest...)
    ltol = 0.25 / sqrt(H.n); SOL_feas = nothing
    need = Ref{Int32}(0); b = Ref(0.0); zabs = Ref(0.0)
    p1() = check(ccall((:mgbx_phase1_init, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Float64}, Ref{Float64}), H.h, need, b, zabs), H.h)
    p1()
    if need[] != 0
        check(ccall((:mgbx_attach_feasibility, LIB), Cint, (Ptr{Cvoid}, Ref{CAmg}), H.h, H.feas), H.h); p1()
        Rbox = max(10.0, 10zabs[]); Rmax = max(feasibility_Rmax, Rbox); first = true
        feasible() = scalars(H, 1).var_max[H.nu+1] < 0
        while true
            printlog("mgb_driver: feasibility phase with bounding box R=", Rbox)
            check(ccall((:mgbx_set_feasibility_box, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), H.h, b[], Rbox), H.h)
            first || check(ccall((:mgbx_reset_feasibility_state, LIB), Cint, (Ptr{Cvoid},), H.h), H.h); first = false
            tfirst = Inf
            stop(tt) = feasible() ? (tfirst = min(tfirst, tt); tt >= 2tfirst) : false
            failed = false
            try
                SOL_feas = core(H, 1; t = t_feasibility, early_stop = stop, stop_lambda_tol = ltol, rest...)
            catch e
                e isa InterruptException && rethrow(); failed = true
            end
            if !failed
                feasible() && break
                sc = scalars(H, 1); vmax = maximum(sc.var_absmax[1:H.nu])
                vmax <= Rbox / 2 && throw(MGB.MGBConvergenceFailure("The problem appears to be infeasible ...", :infeasible))
            end
            10Rbox > Rmax && throw(MGB.MGBConvergenceFailure("Could not find a strictly feasible point ...", :feasibility_Rmax))
            Rbox *= 10
        end
        check(ccall((:mgbx_handoff, LIB), Cint, (Ptr{Cvoid},), H.h), H.h)
        tm = Ref(0.0); tstar = Ref(0.0)
        check(ccall((:mgbx_matched_t, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ref{Float64}, Ref{Float64}), H.h, t, tm, tstar), H.h)
        t = min(t, tm[])
    end
    SOL_main = core(H, 0; t = t, stop_lambda_tol = ltol, rest...)
    z = Matrix{Float64}(undef, H.n, H.nu)
    check(ccall((:mgbx_get_z, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), H.h, 0, z), H.h)
    (z = z, SOL_feasibility = SOL_feas, SOL_main = SOL_main)
end

"Drop-in for mgb_solve(prob) (src/mgb.jl:798-842) on one B200."
function mgb_solve_b200(mg, prob; rest...)
    w = prob.M[1].w; sel = w .!= 0
    bw = all(sel) ? nothing : Float64.(sel) ./ count(sel)                      # src/convex.jl:279-304 with the default mask
    H = Handle(mg, prob; barrier_weights = bw)
    try
        S = driver(H; rest...)
        return MGB.MGBSOL(S.z, S.SOL_feasibility, S.SOL_main, "mgb_solve: device = B200 (libmgbx)", prob.geometry)
    finally
        close!(H)
    end
end

end # module
