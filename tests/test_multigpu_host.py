"""World-size-2 gloo test (CPU) of the sample-partition host logic used for N > 1 (nexus_b200/multigpu.py): disjoint frame
blocks, sum-reduce of accumulation buffers, frame-count bookkeeping."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_frame(frame, n):
    """Stand-in for one rendered frame: a deterministic function of the frame index (as the RNG keying guarantees)."""
    rng = np.random.default_rng(1000 + frame)
    return rng.uniform(0, 2, n).astype(np.float32)


def _worker(rank, world, port, frames_per_rank, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nexus_b200.multigpu import frame_block, reduce_accumulation
    first = frame_block(rank, world, frames_per_rank)
    acc = np.zeros(n, np.float32)
    for f in range(first, first + frames_per_rank):
        acc += _fake_frame(f, n)
    t = torch.from_numpy(acc)
    total = reduce_accumulation(t, frames_per_rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[total], t.numpy()]))
    dist.destroy_process_group()


def test_frame_blocks_are_disjoint_and_cover():
    from nexus_b200.multigpu import frame_block
    for world in (1, 2, 4, 8):
        k = 8
        frames = [f for r in range(world) for f in range(frame_block(r, world, k), frame_block(r, world, k) + k)]
        assert sorted(frames) == list(range(1, world * k + 1))


def test_two_rank_reduce_equals_single_process(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world, k, n = 2, 5, 3 * 16 * 9
    mp.start_processes(_worker, args=(world, port, k, n, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    want = np.zeros(n, np.float64)
    for f in range(1, world * k + 1):
        want += _fake_frame(f, n)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert int(got[0]) == world * k
        assert np.allclose(got[1:], want, rtol=1e-5)
