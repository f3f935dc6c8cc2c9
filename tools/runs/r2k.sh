set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q --durations=8 --timeout=400 --tb=short > gpurun_out/r2k_tests.log 2>&1
timeout 200 python tools/fixture_diff.py fem3d_k1_c24_t0.1 > gpurun_out/r2k_diff_c24_t0.1.json 2>&1
timeout 200 python tools/diag_solve.py q1c64 t=0.01 verbose=0 > gpurun_out/r2k_q1c64.json 2>&1
timeout 500 python tools/bench_parabolic.py 9 > gpurun_out/r2k_parabolic9.json 2> gpurun_out/r2k_parabolic9.err
echo finished
