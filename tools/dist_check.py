"""Multi-GPU check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [case]
Every rank solves its element block of the same problem (NCCL all-reduces inside libmgbx); rank 0 also solves the whole
problem on its own GPU and compares z, the objective history and the Newton counts."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import mgbx
from mgbx import native, solver, geometry as G, hierarchy as H, problem as P

case = sys.argv[1] if len(sys.argv) > 1 else "p1L7"
over = {}
for a in sys.argv[2:]:                      # mgbx_config overrides, e.g. shard_min_rows=100 shard_solve=0
    k, v = a.split("=")
    over[k] = float(v) if ("." in v or "e" in v.lower()) else int(v)
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("gloo")            # host-side plumbing only: the data path uses NCCL inside libmgbx
if case.startswith("p1L"):
    prob = P.assemble(H.amg(G.subdivide(G.fem2d_P1(), int(case[3:]))), p=1.5)
elif case.startswith("q1c"):
    prob = P.assemble(H.amg(G.structured_box(3, int(case[3:]), k=1)), p=1.0)
else:
    raise SystemExit("unknown case")
uid = [native.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
cfg = dict(device=lr, **over)
t0 = time.time()
sol = solver.mgb_solve(prob, comm=(rank, world, uid[0]), config=cfg)
t1 = time.time()
sol2 = solver.mgb_solve(prob, comm=(rank, world, uid[0]), config=cfg) if False else None
parts = [None] * world
dist.all_gather_object(parts, (sol["node_range"], sol["z"]))
its = sol["SOL_main"]["its"]
if rank == 0:
    z = np.concatenate([p[1] for p in sorted(parts, key=lambda q: q[0][0])], axis=0)
    t2 = time.time()
    ref = solver.mgb_solve(prob, config=cfg)
    same_sched = its.shape == ref["SOL_main"]["its"].shape and np.allclose(sol["SOL_main"]["ts"], ref["SOL_main"]["ts"], rtol=1e-12)
    t3 = time.time()
    err = np.linalg.norm(z - ref["z"]) / np.linalg.norm(ref["z"])
    dob = abs(sol["SOL_main"]["c_dot_Dz"][-1] - ref["SOL_main"]["c_dot_Dz"][-1]) / abs(ref["SOL_main"]["c_dot_Dz"][-1])
    same_shape = its.shape == ref["SOL_main"]["its"].shape
    dits = int(np.abs(its.sum(axis=0) - ref["SOL_main"]["its"].sum(axis=0)).max()) if same_shape else -1
    print("DIST %s %s world=%d n=%d pcg_iters %d vs %d: rel z err %.3e, objective rel diff %.3e, t-steps %d vs %d, max Newton-count diff %d, "
          "wall %.2fs (dist, incl. setup) vs %.2fs (1 GPU); stages dist f01 %.0f f2 %.0f solve %.0f ms | single f01 %.0f f2 %.0f solve %.0f ms"
          % (case, over, world, prob.geometry.n, sol["stats"]["pcg_iters"], ref["stats"]["pcg_iters"], err, dob, its.shape[1], ref["SOL_main"]["its"].shape[1], dits, t1 - t0, t3 - t2,
             sol["stats"]["ms_f01"], sol["stats"]["ms_f2"], sol["stats"]["ms_solve"],
             ref["stats"]["ms_f01"], ref["stats"]["ms_f2"], ref["stats"]["ms_solve"]), flush=True)
    assert err < 1e-6 and dob < 1e-8 and same_shape and dits <= 1, "multi-GPU result differs from the single-GPU result"
    print("DIST OK", flush=True)
dist.barrier()
dist.destroy_process_group()
