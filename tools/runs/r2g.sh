set -x
cd $GRAFT_REPO_ROOT
(timeout 2400 python -m pytest tests -m gpu -q --durations=40 2>&1 | tail -120) > gpurun_out/r2g_tests.log 2>&1
timeout 300 python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2g_plain.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 14000 --csv --log-file gpurun_out/r2g_launches.csv python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2g_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg2 -s 60 -c 2 -o gpurun_out/r2g_pcg2 python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2g_ncu_pcg2.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo finished
