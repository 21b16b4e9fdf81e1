"""Sample-partitioned multi-GPU rendering: one process per GPU, scene replicated, rank g renders its own
block of frame indices into its local float SUM buffer, one all-reduce(sum) of float[3*W*H] over NCCL/NVLink at the end.

The reference has no multi-GPU path (process-global __constant__ state, N/Cuda/PathTracer/PathTracer.cu:21-37); frames are
independent given the frame-number-keyed RNG (N/Cuda/Random.cuh:67-73), which is what makes this partition exact: the
reduced image is the same sum of the same per-frame images a single GPU would accumulate, up to float summation order.

torch.distributed is plumbing only: the tensors handed to all_reduce are zero-copy views of the renderer's own buffer.
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
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def accumulation_tensor(path_tracer, device):
    """Zero-copy torch view of a PathTracer's device accumulation SUM buffer (3*W*H floats)."""
    ptr, _ = path_tracer.AccumulationDevice()
    w, h = path_tracer.resolution
    return torch.as_tensor(_DevView(ptr, 3 * w * h), device=device)


def render_partitioned(path_tracer, scene, frames_per_rank, first_frame=1, stream=None, group=None):
    """Renders this rank's block and reduces.  After the call every rank's accumulation holds all world*frames_per_rank frames."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    path_tracer.Render(scene, frames=frames_per_rank, firstFrame=frame_block(rank, world, frames_per_rank, first_frame))
    if world > 1:
        acc = accumulation_tensor(path_tracer, torch.device("cuda", path_tracer.ctx.device))
        if stream is None:
            path_tracer.ctx.synchronize()
            total = reduce_accumulation(acc, frames_per_rank, group)
        else:
            with torch.cuda.stream(stream):
                total = reduce_accumulation(acc, frames_per_rank, group)
        path_tracer.SetAccumulatedFrames(total)
    return frames_per_rank * world
