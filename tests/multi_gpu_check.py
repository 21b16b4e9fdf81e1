"""Multi-GPU correctness of the sample partition, run under torchrun on N >= 2 GPUs (tests/test_gpu_multi.py launches it; also
`python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py` by hand).

 1. N ranks render one block of frames each and all-reduce: the reduced image equals the image ONE GPU accumulates over the same
    frame indices (frames are pure functions of (pixel, frame, bounce) - the RNG keying - so the only difference is float summation
    order: 1e-5 relative on the sums), and the float64 checksum of the reduced buffer equals the sum of the per-rank checksums.
 2. A second render_partitioned call on top of the first adds only its own frames (the ADVICE case: the accumulation already holds
    the global sum on every rank and must not be reduced again).
 3. Every rank ends with the same buffer.
Prints MULTI_GPU_CHECK_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import nexus_b200 as nx
from nexus_b200 import multigpu, scenes


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = nx.Context(local)
    res = (320, 200)
    desc = scenes.with_triangle_data(scenes.instanced_scene(n_blas=6, n_instances=20, nu=20, nv=18, path_length=4))
    scene = multigpu.build_scene_sharded(ctx, desc, res)
    pt = nx.PathTracer(ctx, res)
    per = 2
    acc = multigpu.accumulation_tensor(pt, torch.device("cuda", local))

    total = multigpu.render_partitioned(pt, scene, per, first_frame=1)
    ctx.synchronize(); torch.cuda.synchronize()
    assert total == per * world and pt.GetFrameNumber() == per * world
    multi = pt.ReadAccumulation() * total                       # sums

    # the same frame indices on this GPU alone
    solo = nx.PathTracer(ctx, res)
    solo.Render(scene, frames=per * world, firstFrame=1)
    single = solo.ReadAccumulation() * (per * world)
    scale = np.abs(single).max()
    assert np.abs(multi - single).max() <= 1e-5 * scale + 1e-4 * np.abs(single).mean(), (np.abs(multi - single).max(), scale)
    assert abs(float(multi.astype(np.float64).sum()) - float(single.astype(np.float64).sum())) <= 1e-6 * float(single.astype(np.float64).sum())

    # every rank holds the same reduced buffer
    mine = acc.sum(dtype=torch.float64)
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(lo) == float(hi)

    # a second call adds only its own frames
    total2 = multigpu.render_partitioned(pt, scene, per, first_frame=1 + per * world)
    ctx.synchronize(); torch.cuda.synchronize()
    assert total2 == 2 * per * world
    solo.Render(scene, frames=per * world, firstFrame=1 + per * world)
    single2 = solo.ReadAccumulation() * (2 * per * world)
    multi2 = pt.ReadAccumulation() * total2
    assert np.abs(multi2 - single2).max() <= 2e-5 * np.abs(single2).max() + 1e-4 * np.abs(single2).mean()

    # and on a caller-supplied stream the reduction is ordered behind the render
    pt.ResetFrameNumber()
    side = torch.cuda.Stream()
    multigpu.render_partitioned(pt, scene, per, first_frame=1, stream=side)
    ctx.synchronize(); torch.cuda.synchronize()
    multi3 = pt.ReadAccumulation() * (per * world)
    assert np.abs(multi3 - single).max() <= 1e-5 * scale + 1e-4 * np.abs(single).mean()

    pt.close(); solo.close(); scene.close(); ctx.close()
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
