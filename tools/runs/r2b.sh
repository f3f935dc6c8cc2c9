set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests -m gpu -x -q -k "generation or persistent or golden or fixture_fem2d or forced_pcg or recovered" 2>&1 | tail -30) > gpurun_out/r2b_tests.log 2>&1
# stagnation diagnostics with the failure rule disabled (first-generation kernel, as in round 1)
timeout 300 python tools/diag_solve.py q1c32 t=0.01 persistent=1 pcg_fail_rtol=1.0 > gpurun_out/r2b_diag_q1c32.json 2> gpurun_out/r2b_diag_q1c32.err
timeout 300 python tools/diag_solve.py parabolic6 persistent=1 pcg_fail_rtol=1.0 > gpurun_out/r2b_diag_parabolic6.json 2> gpurun_out/r2b_diag_parabolic6.err
timeout 300 python tools/diag_solve.py q1c32 t=0.01 verbose=0 > gpurun_out/r2b_q1c32_default.json 2>&1
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-same-config > gpurun_out/r2b_bench_gen2.json 2> gpurun_out/r2b_bench_gen2.err
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-same-config --config pcg_rtol=1e-7 > gpurun_out/r2b_bench_gen2_rtol7.json 2> gpurun_out/r2b_bench_gen2_rtol7.err
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-same-config --config persistent=1 > gpurun_out/r2b_bench_gen1.json 2> gpurun_out/r2b_bench_gen1.err
grep -c . gpurun_out/r2b_diag_q1c32.err
echo finished
