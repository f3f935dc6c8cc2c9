set -x
cd $GRAFT_REPO_ROOT
timeout 120 python -m pytest tests/test_gpu_fixtures.py -m gpu -q -rxXs -s --timeout=100 --tb=short -k "fem2d_P1_L10" > gpurun_out/r2o_test_L10.log 2>&1
timeout 110 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-same-config --fem3d-c 0 --no-profile-pass > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo finished
