"""N-GPU check of the sharded scene build (run under torchrun, one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_build.py
Every rank builds the scene twice — sharded (its own BLASes + all-gather) and plainly (everything locally) — and checks that the two
are the same scene: canonical BLAS bytes, closest hits bit for bit.  Also times both builds."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import nexus_b200 as nx
import oracle_lib as O
from nexus_b200 import multigpu, scenes

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = nx.Context(local)
n_blas = int(sys.argv[1]) if len(sys.argv) > 1 else 64
res = (320, 240)
desc = scenes.instanced_scene(n_blas=n_blas, n_instances=n_blas, nu=40, nv=36, path_length=3)
t0 = time.time(); shard = multigpu.build_scene_sharded(ctx, desc, res); ctx.synchronize(); t_shard = time.time() - t0
t0 = time.time(); plain = scenes.build(ctx, desc, res); ctx.synchronize(); t_plain = time.time() - t0
for k in range(len(desc["meshes"])):
    a, b = plain.MeshBVH(k), shard.MeshBVH(k)
    (na, pa), (nb, pb) = O.canon_bvh8(*a.ToHost()), O.canon_bvh8(*b.ToHost())
    assert a.nodeCount == b.nodeCount and (na == nb).all() and (pa == pb).all() and (a.bounds == b.bounds).all(), (rank, k)
o, d = scenes.camera_rays(desc["camera"], res)
rays = nx.make_rays(o, d)
ha, hb = plain.TraceClosest(rays), shard.TraceClosest(rays)
assert ha.tobytes() == hb.tobytes(), rank
hit = int((ha["prim"] != 0xffffffff).sum())
if world > 1:
    t = torch.tensor([t_shard, t_plain], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_shard, t_plain = float(t[0]), float(t[1])
    dist.barrier()
if rank == 0:
    print(f"sharded build OK on {world} rank(s): {len(desc['meshes'])} BLASes identical, {len(rays)} closest hits identical ({hit} hit); "
          f"scene build {t_shard:.2f} s sharded vs {t_plain:.2f} s replicated", flush=True)
plain.close(); shard.close(); ctx.close()
if world > 1:
    dist.destroy_process_group()
