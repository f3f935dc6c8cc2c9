set -x
cd $GRAFT_REPO_ROOT
timeout 560 python -m pytest tests -m gpu -q --durations=6 --timeout=300 --tb=short > gpurun_out/r2l_tests.log 2>&1
timeout 400 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo finished
