set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
# correctness of the row-sharded solve: small problems with sharding forced, against the single-GPU solve
timeout 300 $TR --master-port 29511 tools/dist_check.py p1L6 shard_min_rows=100 dense_direct_max=64 coarse_max=64 > gpurun_out/r2f_dist_p1L6_sh.log 2>&1
timeout 300 $TR --master-port 29512 tools/dist_check.py q1c8 shard_min_rows=100 dense_direct_max=64 coarse_max=64 > gpurun_out/r2f_dist_q1c8_sh.log 2>&1
timeout 300 $TR --master-port 29513 tools/dist_check.py p1L8 shard_min_rows=1000 > gpurun_out/r2f_dist_p1L8_sh.log 2>&1
timeout 300 $TR --master-port 29514 tools/dist_check.py p1L8 shard_solve=0 > gpurun_out/r2f_dist_p1L8_rep.log 2>&1
# C2 on 2 GPUs: sharded (default) and replicated solve
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2f_bench2_sh.json 2> gpurun_out/r2f_bench2_sh.err
timeout 600 $TR --master-port 29516 bench.py --gpus 2 --steps 3 --warmup 2 --config shard_solve=0 > gpurun_out/r2f_bench2_rep.json 2> gpurun_out/r2f_bench2_rep.err
# fem3d 64^3 (2.1 M DOF) on 2 GPUs: sharded and replicated solve
timeout 600 $TR --master-port 29517 tools/dist_fem3d.py 64 0.01 > gpurun_out/r2f_q1c64_sh.jsonl 2> gpurun_out/r2f_q1c64_sh.err
timeout 600 $TR --master-port 29518 tools/dist_fem3d.py 64 0.01 shard_solve=0 > gpurun_out/r2f_q1c64_rep.jsonl 2> gpurun_out/r2f_q1c64_rep.err
echo finished
