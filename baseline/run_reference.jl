#!/usr/bin/env julia
#
# run_reference.jl -- time the UNMODIFIED reference (MultiGridBarrier.jl v1.1.0) on bench.py's workload family and dump
# what the north-star's gates compare: z, the objective history, Newton counts, the t-schedule.
#
# STATUS: NOT executed in this repository's build image (no Julia there; BASELINE.md section 3).  It is the script for a
# box that has Julia; its outputs are what tests/golden/size_*.npz hold for the CPU oracle, so the two can be diffed.
# Follows the timing pattern of the reference's own tool, tools/bench_cuda_vs_native.jl:19-24,121-130
# (`@timed mgb_solve(prob; device=CPUDevice, verbose=false)` after one GC, problem assembled once outside the timer).
#
#   julia --project=/path/to/MultiGridBarrier.jl baseline/run_reference.jl
#
# Environment variables (same names as the reference tool where they overlap):
#   BENCH_FEM   "fem2d_P1" (default: config C2's family) | "fem2d_P2" | "fem3d"
#   BENCH_L     subdivision level (default 8: n = 98 304, the level bench.py --impl reference runs; 10 = the headline)
#   BENCH_P     p-Laplace exponent (default 1.5)
#   BENCH_CUDA  "1": also time device=CUDADevice (needs CUDA.jl + CUDSS_jll), listed for context as the north-star asks
#   BENCH_OUT   output prefix (default "reference_<fem>_L<L>_p<p>"): writes <prefix>.json and <prefix>_z.bin (Float64, column-major)
using MultiGridBarrier
using Printf

const FEM = get(ENV, "BENCH_FEM", "fem2d_P1")
const L = parse(Int, get(ENV, "BENCH_L", "8"))
const P = parse(Float64, get(ENV, "BENCH_P", "1.5"))
const WITH_CUDA = get(ENV, "BENCH_CUDA", "0") == "1"
const OUT = get(ENV, "BENCH_OUT", @sprintf("reference_%s_L%d_p%g", FEM, L, P))

geom = if FEM == "fem2d_P1"
    subdivide(fem2d_P1(), L)
elseif FEM == "fem2d_P2"
    subdivide(fem2d_P2(), L)
elseif FEM == "fem3d"
    subdivide(fem3d(; k = 1), L)
else
    error("unknown BENCH_FEM=$FEM")
end
n = size(geom.x, 1) * size(geom.x, 2)            # broken nodes V*N (the DOF count of bench.md and of bench.py)
t_setup = @elapsed (prob = assemble(amg(geom); p = P))
@printf("%s L=%d p=%g: n=%d broken nodes, setup (amg + assemble) %.2fs, %d threads\n", FEM, L, P, n, t_setup, Threads.nthreads())

mgb_solve(assemble(amg(subdivide(fem2d_P1(), 2)); p = P); verbose = false)   # compile
GC.gc(true)
b = @timed mgb_solve(prob; device = CPUDevice, verbose = false)
sol = b.value
its = sol.SOL_main.its
newton = sum(its)
@printf("CPU: %.3fs, %d Newton steps, %.4g DOF*Newton-steps/s, objective %.17g\n", b.time, newton, n * newton / b.time,
        sol.SOL_main.c_dot_Dz[end])

t_cuda = NaN
if WITH_CUDA
    @eval using CUDA, CUDSS_jll
    mgb_solve(assemble(amg(subdivide(fem2d_P1(), 2)); p = P); device = CUDADevice, verbose = false)
    GC.gc(true); CUDA.reclaim()
    bc = @timed mgb_solve(prob; device = CUDADevice, verbose = false)
    CUDA.synchronize()
    t_cuda = bc.time
    @printf("CUDAExt: %.3fs (max |z - z_cpu| = %.2e)\n", t_cuda, maximum(abs.(bc.value.z .- sol.z)))
end

open(OUT * "_z.bin", "w") do io
    write(io, Float64.(sol.z))
end
open(OUT * ".json", "w") do io
    fmt(v) = "[" * join(string.(v), ", ") * "]"
    println(io, "{")
    println(io, "  \"fem\": \"$FEM\", \"L\": $L, \"p\": $P, \"n\": $n, \"threads\": $(Threads.nthreads()),")
    println(io, "  \"setup_s\": $t_setup, \"cpu_solve_s\": $(b.time), \"cuda_solve_s\": $(isnan(t_cuda) ? "null" : t_cuda),")
    println(io, "  \"newton_steps\": $newton, \"value_dof_newton_steps_per_s\": $(n * newton / b.time),")
    println(io, "  \"its_per_step\": $(fmt(vec(sum(its; dims = 1)))),")
    println(io, "  \"ts\": $(fmt(sol.SOL_main.ts)),")
    println(io, "  \"c_dot_Dz\": $(fmt(sol.SOL_main.c_dot_Dz)),")
    println(io, "  \"z_shape\": $(fmt(collect(size(sol.z)))), \"z_file\": \"$(OUT)_z.bin\"")
    println(io, "}")
end
println("wrote $(OUT).json and $(OUT)_z.bin")
