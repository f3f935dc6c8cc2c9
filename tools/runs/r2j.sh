set -x
cd $GRAFT_REPO_ROOT
(timeout 1500 python -m pytest tests -m gpu -q --durations=25 --timeout=600 2>&1 | tail -80) > gpurun_out/r2j_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err
echo finished
