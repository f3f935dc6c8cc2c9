set -x
cd $GRAFT_REPO_ROOT
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 tools/dist_check.py p1L6 shard_min_rows=100 shard_min_nnz=0 dense_direct_max=64 coarse_max=64 > gpurun_out/r2i_dist_p1L6_sh_n$N.log 2>&1
timeout 200 $TR --master-port 29522 tools/dist_check.py q1c8 shard_min_rows=100 shard_min_nnz=0 dense_direct_max=64 coarse_max=64 > gpurun_out/r2i_dist_q1c8_sh_n$N.log 2>&1
timeout 600 $TR --master-port 29523 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err
echo finished
