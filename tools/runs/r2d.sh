set -x
cd $GRAFT_REPO_ROOT
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2d_tests.log 2>&1
timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-same-config --config elem_bulk=0 > gpurun_out/r2d_bench_nobulk.json 2> gpurun_out/r2d_bench_nobulk.err
MGBX_PCG_PROF=1 timeout 300 python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2d_prof_c2.json 2> gpurun_out/r2d_prof_c2.err
timeout 300 python tools/diag_solve.py q1c32 t=0.01 verbose=0 > gpurun_out/r2d_q1c32.json 2>&1
timeout 300 python tools/diag_solve.py q1c32 t=0.01 verbose=0 persistent=1 > gpurun_out/r2d_q1c32_gen1.json 2>&1
timeout 600 python tools/diag_solve.py q1c64 t=0.01 verbose=0 > gpurun_out/r2d_q1c64.json 2>&1
echo finished
