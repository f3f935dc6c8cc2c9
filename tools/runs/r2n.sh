set -x
cd $GRAFT_REPO_ROOT
timeout 150 python -m pytest tests -m gpu -q -x --timeout=100 --tb=short -k "golden or stages or bulk or recovered or fem2d_P1_L7 or fem3d_k1_c16 or spectral2d_n9 or feasib or midsize" > gpurun_out/r2n_tests.log 2>&1
timeout 150 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-same-config --fem3d-c 0 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo finished
