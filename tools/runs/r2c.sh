set -x
cd $GRAFT_REPO_ROOT
# phase profile of the gen-2 kernel on C2 (one solve)
MGBX_PCG_PROF=1 timeout 300 python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2c_prof_c2.json 2> gpurun_out/r2c_prof_c2.err
# stagnation experiments, 3-D
for w in 25 100 400; do
  timeout 300 python tools/diag_solve.py q1c32 t=0.01 verbose=0 pcg_stall_window=$w direct_fallback=0 > gpurun_out/r2c_q1c32_w$w.json 2>&1
done
timeout 300 python tools/diag_solve.py q1c32 t=0.01 verbose=0 pcg_stall_window=100 pcg_rtol=1e-7 lambda_power=6 direct_fallback=0 > gpurun_out/r2c_q1c32_w100_lp6.json 2>&1
# parabolic L6 with the direct fallback (default) and a wide window
timeout 400 python tools/diag_solve.py parabolic6 verbose=1 > gpurun_out/r2c_parabolic6.json 2> gpurun_out/r2c_parabolic6.err
timeout 400 python tools/diag_solve.py parabolic6 verbose=0 pcg_stall_window=200 direct_fallback=0 > gpurun_out/r2c_parabolic6_w200.json 2>&1
# forced PCG on pure P2 (600 unknowns), diagnostics
timeout 200 python tools/diag_solve.py purep2_4 verbose=1 dense_direct_max=0 coarse_max=0 direct_fallback=0 > gpurun_out/r2c_purep2_forced.json 2> gpurun_out/r2c_purep2_forced.err
timeout 200 python tools/diag_solve.py purep2_6 verbose=0 > gpurun_out/r2c_purep2_6.json 2>&1
# C2 with the coarse-level Newton systems on PCG instead of the dense factor
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-same-config --config pcg_rtol=1e-7 --config dense_direct_max=512 > gpurun_out/r2c_bench_ddm512.json 2> gpurun_out/r2c_bench_ddm512.err
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2c_tests.log 2>&1
echo finished
