"""C4 on N GPUs: fem3d k=1 on c^3 hexahedra, p = 1, element-partitioned (one process per GPU).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/dist_fem3d.py 100 0.01"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P
c = int(sys.argv[1]); t_init = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
over = {}
for a in sys.argv[3:]:                      # mgbx_config overrides, e.g. shard_solve=0
    k, v = a.split("=")
    over[k] = float(v) if ("." in v or "e" in v.lower()) else int(v)
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
t0 = time.time()
prob = P.assemble(H.amg(G.structured_box(3, c, k=1)), p=1.0)
tb = time.time() - t0
uid = [native.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
dist.barrier()
t0 = time.time()
sol = solver.mgb_solve(prob, comm=(rank, world, uid[0]) if world > 1 else None, config=dict(device=lr, **over), t=t_init)
dt = time.time() - t0
st = sol["stats"]
mem = torch.cuda.mem_get_info()
out = dict(config=over, rank=rank, world=world, nodes=prob.geometry.n, elements=prob.geometry.N, host_build_s=round(tb, 1), solve_wall_s=round(dt, 1),
           create_s=round(st["create_s"], 1), newton_steps=int(sol["SOL_main"]["its"].sum()), barrier_steps=int(sol["SOL_main"]["its"].shape[1]),
           pcg_iters=st["pcg_iters"], ms_f01=round(st["ms_f01"]), ms_f2=round(st["ms_f2"]), ms_solve=round(st["ms_solve"]),
           objective=float(sol["SOL_main"]["c_dot_Dz"][-1]), gpu_mem_used_gb=round((mem[1] - mem[0]) / 1e9, 1),
           local_nodes=int(sol["z"].shape[0]))
allout = [None] * world
dist.all_gather_object(allout, out)
if rank == 0:
    for o in allout:
        print(json.dumps(o), flush=True)
dist.barrier()
dist.destroy_process_group()
