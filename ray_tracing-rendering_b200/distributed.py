"""Multi-GPU plumbing: one process per GPU, torch.distributed for the single exchange.

The reference parallelises over 16x16 image tiles on CPU threads (renderer.h:40-94).
Every path sample is independent, so the GPUs split the SAMPLES of every pixel
(rank r renders the samples s with s % world == r) and, when a job has fewer samples than
ranks, the image ROWS (plan_split); the scene is replicated; the only
communication is one SUM-reduce of the float4 accumulators at the end (NCCL over
NVLink on GPUs, gloo in the CPU tests).  The per-sample RNG stream depends only on
(pixel, s, seed), so the reduced image equals the single-GPU image up to float
summation order.
"""
import numpy as np


def rank_split(spp: int, rank: int, world: int):
    """(sample_offset, sample_stride, local_spp) of `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    local = (spp - rank + world - 1) // world if rank < spp else 0
    return rank, world, local


def plan_split(spp: int, height: int, rank: int, world: int):
    """How `world` GPUs share one job: {"sample_offset", "sample_stride", "row_offset", "row_stride"}.

    The samples of every pixel are split when there are enough of them (spp >= world): every
    rank then touches every pixel and the ranks stay balanced whatever the image shows.  With
    fewer samples than ranks (previews) the image ROWS are interleaved across groups of ranks
    instead (rank r renders rows j % row_stride == row_offset) and the samples are split inside
    each group — the GPU counterpart of the reference's tile queue (renderer.h:40-94)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    sample_ways = max(1, min(world, spp))
    while world % sample_ways:
        sample_ways -= 1
    row_ways = world // sample_ways
    if row_ways > max(height, 1):
        raise ValueError("more ranks than image rows x samples")
    return {"sample_offset": rank % sample_ways, "sample_stride": sample_ways,
            "row_offset": rank // sample_ways, "row_stride": row_ways}


def local_rows(height: int, row_offset: int, row_stride: int):
    return np.arange(row_offset, height, max(row_stride, 1))


def local_sample_indices(spp: int, rank: int, world: int):
    off, stride, n = rank_split(spp, rank, world)
    return np.arange(n) * stride + off


def reduce_sum(tensor, dst: int = 0, group=None):
    """SUM-reduce the accumulators onto rank `dst` (in place).  No-op for world size 1."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def render_distributed(ctx, params_fn, spp, width, height, device):
    """Each rank renders its share (plan_split) into a device tensor, then one reduce.
    params_fn(sample_offset=, sample_stride=, row_offset=, row_stride=) -> RenderParams.
    Returns (tensor, stats)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    accum = torch.empty((height, width, 4), dtype=torch.float32, device=device)
    stats = ctx.render_device(params_fn(**plan_split(spp, height, rank, world)), accum.data_ptr(),
                              torch.cuda.current_stream(device).cuda_stream)
    reduce_sum(accum)
    return accum, stats
