"""Sample-partitioned multi-GPU rendering: one process per GPU, scene replicated, rank g renders its own
block of frame indices into its local float SUM buffer, one all-reduce(sum) of float[3*W*H] over NCCL/NVLink at the end.

The reference has no multi-GPU path (process-global __constant__ state, N/Cuda/PathTracer/PathTracer.cu:21-37); frames are
independent given the frame-number-keyed RNG (N/Cuda/Random.cuh:67-73), which is what makes this partition exact: the
reduced image is the same sum of the same per-frame images a single GPU would accumulate, up to float summation order.

torch.distributed is plumbing only: the tensors handed to all_reduce are zero-copy views of the renderer's own buffer.

Sharded scene build ("BVH build, many meshes"): the BLAS of a mesh is an independent unit of work, so rank g
builds the BLASes of meshes g, g + G, g + 2G, ... and one all-gather hands every rank every BLAS; the TLAS (one primitive per
instance) is built redundantly by each rank.  The imported trees are the bytes the owner built, so every rank ends up with
the scene a single-GPU build produces.
"""
import torch
import torch.distributed as dist


def frame_block(rank, world, frames_per_rank, first_frame=1):
    """First frame index of rank's block: blocks are contiguous and disjoint, together [first, first + world*frames)."""
    if not (0 <= rank < world) or frames_per_rank < 0:
        raise ValueError("bad partition")
    return first_frame + rank * frames_per_rank


def reduce_accumulation(acc_sum, local_frames, group=None):
    """In-place all-reduce(sum) of the accumulation SUM buffers; returns the total number of frames now in the buffer.
    acc_sum: 1-D float32 tensor (CUDA for NCCL, CPU for gloo)."""
    if acc_sum.dtype != torch.float32:
        raise TypeError("accumulation buffers are float32")
    count = torch.tensor([float(local_frames)], dtype=torch.float64, device=acc_sum.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc_sum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    return int(round(float(count[0])))


class _DevView:
    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def accumulation_tensor(path_tracer, device):
    """Zero-copy torch view of a PathTracer's device accumulation SUM buffer (3*W*H floats)."""
    ptr, _ = path_tracer.AccumulationDevice()
    w, h = path_tracer.resolution
    return torch.as_tensor(_DevView(ptr, 3 * w * h), device=device)


def render_partitioned(path_tracer, scene, frames_per_rank, first_frame=1, stream=None, group=None):
    """Renders this rank's block of frames and reduces.  After the call every rank's accumulation holds everything it held before
    (which must be the same on every rank: nothing, or the result of earlier calls) plus all world * frames_per_rank new frames.

    Only THIS call's contribution is all-reduced: an accumulation that already holds frames is the global sum on every rank, and
    reducing it again would count it world times.  The contribution is taken as a device-side difference against a snapshot.
    The reduction is ordered behind the render on the device: it runs on the context's own stream, or - when `stream` is given -
    on that stream after an event recorded on the context's stream (nx_renderer_render leaves it ordered behind the shadow-ray
    stream, so the accumulation is complete at that point)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    before = path_tracer.GetFrameNumber()
    if world == 1:
        path_tracer.Render(scene, frames=frames_per_rank, firstFrame=frame_block(rank, world, frames_per_rank, first_frame))
        return before + frames_per_rank
    device = torch.device("cuda", path_tracer.ctx.device)
    from ._capi import lib
    render_stream = torch.cuda.ExternalStream(int(lib().nx_ctx_stream(path_tracer.ctx._h)), device=device)
    acc = accumulation_tensor(path_tracer, device)
    snapshot = None
    if before > 0:
        with torch.cuda.stream(render_stream):
            snapshot = acc.clone()
    path_tracer.Render(scene, frames=frames_per_rank, firstFrame=frame_block(rank, world, frames_per_rank, first_frame))
    run_on = render_stream
    if stream is not None:
        stream.wait_event(render_stream.record_event())
        run_on = stream
    with torch.cuda.stream(run_on):
        if snapshot is None:
            added = reduce_accumulation(acc, frames_per_rank, group)
        else:
            acc.sub_(snapshot)
            added = reduce_accumulation(acc, frames_per_rank, group)
            acc.add_(snapshot)
    if stream is not None:
        render_stream.wait_event(stream.record_event())       # later renders and reads on the context's stream see the reduced buffer
    path_tracer.SetAccumulatedFrames(before + added)
    return before + added


# ------------------------------------------------------------------------------------------ sharded BLAS builds ----
NODE_WORDS = 20      # one CWBVH8 node = 80 bytes = 20 x int32


def mesh_owner(mesh_idx, world):
    """Round-robin: mesh k belongs to rank k mod world (meshes of a scene are similar in size; no balancing beyond that)."""
    return mesh_idx % world


def blas_layout(node_counts, prim_counts, world):
    """Where every mesh's BLAS sits in its owner's packed payload.  Returns (offsets, sizes): offsets[k] = (word offset of the
    nodes, word offset of the primitive indices) inside the payload of rank mesh_owner(k), sizes[g] = payload words of rank g.
    Payload of a rank: for each of its meshes in index order, NODE_WORDS * nodes words of nodes, then one word per primitive."""
    offsets, sizes = [], [0] * world
    for k, (nn, pn) in enumerate(zip(node_counts, prim_counts)):
        g = mesh_owner(k, world)
        offsets.append((sizes[g], sizes[g] + NODE_WORDS * int(nn)))
        sizes[g] += NODE_WORDS * int(nn) + int(pn)
    return offsets, sizes


def exchange_blas(local, n_meshes, device, group=None):
    """All-gather of the BLASes every rank built.  local: {mesh index: (nodes int32[NODE_WORDS * N], prim_idx int32[n],
    bounds float32[6])} for exactly the meshes this rank owns (tensors on `device`).  Returns {mesh index: (nodes, prim_idx,
    bounds)} for ALL meshes as views into one gathered buffer (CUDA + NCCL in production, CPU + gloo in the tests).
    Two collectives: an all-reduce(sum) of the int32 count / bounds table (every row is non-zero on exactly one rank, so the sum
    is exact and carries the float bounds as bit patterns), then one all-gather of the packed payloads padded to the largest."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    owned = [k for k in range(n_meshes) if mesh_owner(k, world) == rank]
    if sorted(local) != owned:
        raise ValueError(f"rank {rank} must supply exactly its own meshes {owned[:4]}..., got {sorted(local)[:4]}...")
    table = torch.zeros((n_meshes, 8), dtype=torch.int32)                    # node count, primitive count, bounds bits
    for k in owned:
        nodes, prim, bounds = local[k]
        if nodes.dtype != torch.int32 or prim.dtype != torch.int32 or nodes.numel() % NODE_WORDS:
            raise TypeError("BLAS payloads are int32 words (NODE_WORDS per node)")
        table[k, 0] = nodes.numel() // NODE_WORDS
        table[k, 1] = prim.numel()
        table[k, 2:8] = bounds.to(device="cpu", dtype=torch.float32).contiguous().view(torch.int32)
    table = table.to(device)
    if world > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    host = table.cpu()
    offsets, sizes = blas_layout(host[:, 0].tolist(), host[:, 1].tolist(), world)
    width = max(max(sizes), 1)
    mine = torch.zeros(width, dtype=torch.int32, device=device)
    for k in owned:
        nodes, prim, _ = local[k]
        a, b = offsets[k]
        mine[a:a + nodes.numel()] = nodes
        mine[b:b + prim.numel()] = prim
    if world > 1:
        parts = [torch.empty(width, dtype=torch.int32, device=device) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
    else:
        parts = [mine]
    out = {}
    for k in range(n_meshes):
        a, b = offsets[k]
        buf = parts[mesh_owner(k, world)]
        out[k] = (buf[a:a + NODE_WORDS * int(host[k, 0])], buf[b:b + int(host[k, 1])], host[k, 2:8].clone().view(torch.float32))
    return out


def build_scene_sharded(ctx, desc, resolution, group=None):
    """scenes.build with the BLAS builds partitioned over the ranks of `group` (NCCL).  Every rank returns the full scene."""
    import nexus_b200 as nx
    from nexus_b200 import scenes
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    device = torch.device("cuda", ctx.device)
    local, handles = {}, []
    for k, m in enumerate(desc["meshes"]):
        if mesh_owner(k, world) != rank:
            continue
        bvh = nx.BuildBLAS(ctx, m["triangles"])          # synchronises the context's stream
        handles.append(bvh)                              # zero-copy views of the builder's outputs: packed straight into the payload
        local[k] = (torch.as_tensor(_DevView(bvh.h.nodes, NODE_WORDS * bvh.nodeCount, "<i4"), device=device),
                    torch.as_tensor(_DevView(bvh.h.prim_idx, bvh.primCount, "<i4"), device=device), torch.from_numpy(bvh.bounds.copy()))
    every = exchange_blas(local, len(desc["meshes"]), device, group)
    torch.cuda.synchronize(device)                       # the gathered buffer is complete before the context's stream copies from it
    for bvh in handles:
        bvh.Free()
    blas = {k: (n.data_ptr(), n.numel() // NODE_WORDS, p.data_ptr(), b.numpy()) for k, (n, p, b) in every.items()}
    scene = scenes.build(ctx, desc, resolution, blas=blas)
    ctx.synchronize()                                    # imports have copied out of the gathered buffer before it is released
    return scene
