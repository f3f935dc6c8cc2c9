set -x
cd $GRAFT_REPO_ROOT
# whole GPU suite, no -x, with per-test durations
(timeout 2400 python -m pytest tests -m gpu -q --durations=50 2>&1 | tail -150) > gpurun_out/r2e_tests.log 2>&1
# ncu: launch list of one C2 solve and a full capture of the gen-2 solve kernel + the element kernels (same command, already run plain in r2d)
timeout 300 python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2e_plain.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 6000 --csv --log-file gpurun_out/r2e_launches.csv python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2e_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg2 -s 40 -c 2 -o gpurun_out/r2e_pcg2 python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2e_ncu_pcg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elem_plap -s 200 -c 4 -o gpurun_out/r2e_elem python tools/diag_solve.py 10 verbose=0 > gpurun_out/r2e_ncu_elem.log 2>&1
echo finished
