set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2a_tests.log 2>&1
timeout 120 ./tools/micro/bench_spmv_phase > gpurun_out/r2a_spmv_phase.log 2>&1
timeout 60 ./tools/micro/bench_barrier > gpurun_out/r2a_barrier.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 400 python tools/tune_smoother.py 10 tools/rtol_sweep.json > gpurun_out/r2a_rtol_sweep.jsonl 2>&1
timeout 600 python tools/tune_smoother.py q1c32 tools/lambda_sweep.json > gpurun_out/r2a_lambda_sweep.jsonl 2>&1
echo finished
