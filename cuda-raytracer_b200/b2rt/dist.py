"""Multi-GPU plumbing: sample sharding + the single accumulation reduce (torch.distributed).

The reference is single-GPU (no NCCL, no streams).  Here every path is independent, so the
samples of every pixel are dealt round-robin to the ranks (rank r renders global samples r, r+G, r+2G, ...),
the scene is replicated, and the per-GPU accumulation buffers (float4 per pixel: rgb sum + sample count)
are combined with ONE reduce per frame.  RNG streams are keyed on the GLOBAL sample index, so the image is
the same for any G up to fp32 summation order.
"""


def shard_samples(total_spp, rank, world):
    """-> (sample_first, sample_stride, local_count) for this rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    local = total_spp // world + (1 if rank < total_spp % world else 0)
    return rank, world, local


def reduce_accum(accum, dst=0, group=None):
    """Sum the per-rank accumulation buffers onto `dst` (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def resolve_mean(accum):
    """accum [h*w*4] (rgb sum, count) -> mean rgb [h*w, 3] (what k_resolve_image computes on the device)."""
    a = accum.reshape(-1, 4)
    cnt = a[:, 3:4].clamp(min=1.0)
    return a[:, :3] / cnt
