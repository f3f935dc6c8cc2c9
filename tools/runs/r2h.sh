set -x
cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k spectral --durations=10 --timeout=300 2>&1 | tail -40) > gpurun_out/r2h_tests_spectral.log 2>&1
(timeout 1200 python -m pytest tests/test_gpu_fixtures.py -m gpu -q -k "spectral or pure_p2 or high_order or parabolic or recovered" --durations=20 --timeout=400 2>&1 | tail -60) > gpurun_out/r2h_tests_fixtures.log 2>&1
timeout 400 python tools/bench_spectral.py 32 64 > gpurun_out/r2h_spectral_kron.jsonl 2> gpurun_out/r2h_spectral_kron.err
timeout 300 python tools/bench_spectral.py 32 spectral_kron=0 > gpurun_out/r2h_spectral_nokron.jsonl 2> gpurun_out/r2h_spectral_nokron.err
timeout 900 python tools/bench_parabolic.py 9 > gpurun_out/r2h_parabolic9.json 2> gpurun_out/r2h_parabolic9.err
MGBX_PCG_PROF=1 timeout 400 python tools/diag_solve.py q1c64 t=0.01 verbose=0 > gpurun_out/r2h_prof_q1c64.json 2> gpurun_out/r2h_prof_q1c64.err
timeout 300 python tools/fixture_diff.py fem3d_k1_c24_t0.1 > gpurun_out/r2h_diff_c24_t0.1.json 2>&1
timeout 300 python tools/fixture_diff.py fem3d_k1_c24_t0.01 > gpurun_out/r2h_diff_c24_t0.01.json 2>&1
timeout 300 python tools/fixture_diff.py fem3d_k1_c24_t0.1 pcg_rtol=1e-9 > gpurun_out/r2h_diff_c24_t0.1_rtol9.json 2>&1
echo finished
