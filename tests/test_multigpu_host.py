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


# ---------------------------------------------------------------------------------------- sharded BLAS builds ----
def _fake_blas(k):
    """Stand-in for the BLAS of mesh k: deterministic, mesh-dependent sizes and bit patterns (negative words, -0.0 bounds)."""
    rng = np.random.default_rng(77 + k)
    nodes = rng.integers(-2**31, 2**31 - 1, 20 * (1 + k % 5), dtype=np.int64).astype(np.int32)
    prim = rng.permutation(3 + 7 * (k % 4)).astype(np.int32)
    bounds = rng.uniform(-5, 5, 6).astype(np.float32)
    if k == 1:
        bounds[0] = -0.0
    return nodes, prim, bounds


def _blas_worker(rank, world, port, n_meshes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nexus_b200.multigpu import exchange_blas, mesh_owner
    local = {k: tuple(torch.from_numpy(a) for a in _fake_blas(k)) for k in range(n_meshes) if mesh_owner(k, world) == rank}
    got = exchange_blas(local, n_meshes, torch.device("cpu"))
    np.savez(os.path.join(out_dir, f"b{rank}.npz"), **{f"{k}_{j}": t.numpy() for k, v in got.items() for j, t in enumerate(v)})
    bad = False
    try:                                                   # a rank that supplies somebody else's mesh is refused
        exchange_blas({k: v for k, v in local.items()} | {(rank + 1) % world: local[next(iter(local))]}, n_meshes, torch.device("cpu"))
    except ValueError:
        bad = True
    assert bad
    dist.destroy_process_group()


def test_blas_layout_partitions_every_payload():
    from nexus_b200.multigpu import NODE_WORDS, blas_layout, mesh_owner
    nodes, prims = [3, 1, 4, 1, 5, 9, 2], [10, 20, 30, 40, 50, 60, 70]
    for world in (1, 2, 3, 8):
        offsets, sizes = blas_layout(nodes, prims, world)
        assert len(sizes) == world and sum(sizes) == NODE_WORDS * sum(nodes) + sum(prims)
        for g in range(world):                             # the slices of a rank's meshes tile its payload exactly
            spans = sorted((a, a + NODE_WORDS * nodes[k]) for k, (a, b) in enumerate(offsets) if mesh_owner(k, world) == g) + \
                    sorted((b, b + prims[k]) for k, (a, b) in enumerate(offsets) if mesh_owner(k, world) == g)
            covered = sorted(spans)
            assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
            assert (covered[0][0] == 0 and covered[-1][1] == sizes[g]) if covered else sizes[g] == 0


def test_two_rank_blas_exchange_delivers_every_mesh_bit_for_bit(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world, n_meshes = 2, 7
    mp.start_processes(_blas_worker, args=(world, port, n_meshes, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    for r in range(world):
        got = np.load(tmp_path / f"b{r}.npz")
        for k in range(n_meshes):
            for j, want in enumerate(_fake_blas(k)):
                assert got[f"{k}_{j}"].dtype == want.dtype and got[f"{k}_{j}"].tobytes() == want.tobytes(), (r, k, j)


def test_single_process_blas_exchange_is_the_identity():
    from nexus_b200.multigpu import exchange_blas
    local = {k: tuple(torch.from_numpy(a) for a in _fake_blas(k)) for k in range(4)}
    got = exchange_blas(local, 4, torch.device("cpu"))
    for k in range(4):
        for j, want in enumerate(_fake_blas(k)):
            assert got[k][j].numpy().tobytes() == want.tobytes()
