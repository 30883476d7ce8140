"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The hot path partitions by independent sequences: rank r runs sequence r on its own GPU
and no data-path collective exists (mode 1).  The only cross-rank traffic is
timing (max over ranks).  For the optional point-sharded mode of very dense scans (mode
2) every rank linearises its shard of the correspondences and the tiny per-pair
normal-equation blocks are summed: one all-reduce of 91*P doubles (error: P doubles)."""
from __future__ import annotations

import numpy as np


def sequence_of_rank(rank: int, world: int, base_sequence: int = 0) -> int:
    """Sequence id processed by a rank (weak scaling: one sequence per GPU)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    return base_sequence + rank


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of n keypoints / correspondences for a rank."""
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(seconds: float, device=None) -> float:
    """Timing rule: the job takes as long as its slowest rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def allreduce_blocks(blocks: np.ndarray, device=None) -> np.ndarray:
    """Sum per-pair blocks (P x 91, or P errors) over the ranks of a point-sharded job.
    The blocks are additive in the correspondences (A^T A), so the sum of the shard blocks
    is the block of the whole pair."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return blocks
    t = torch.from_numpy(np.ascontiguousarray(blocks, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def sharded_linearize(ctx, pairs: np.ndarray, poses: np.ndarray, error_only: bool = False):
    """Point-sharded stage 3 on the CUDA path (SURVEY 8e mode 2).  Every rank holds a replica
    of the sequence's context (`ctx`, created on torch's current NON-default stream -
    `with torch.cuda.stream(torch.cuda.Stream()): Context(params, stream=...cuda_stream)` - and
    switched to its shard with ctx.set_shard(rank, world)); each linearises its share of every pair into
    device memory and ONE NCCL all-reduce of 91 * P doubles (errors: P) over NVLink sums the
    blocks - queued on the same stream, so there is no host round trip before the
    collective.  Returns the full blocks (P x 91) / errors (P) on the host."""
    import torch
    import torch.distributed as dist

    n = int(pairs.shape[0])
    cur = torch.cuda.current_stream()
    same_stream = ctx.stream_handle is not None and ctx.stream_handle == cur.cuda_stream
    out = torch.empty(n * (1 if error_only else 91), dtype=torch.float64, device="cuda")
    if not same_stream:
        cur.synchronize()  # `out` must exist before a foreign stream writes it
    if error_only:
        ctx.error_device(pairs, poses, out.data_ptr())
    else:
        ctx.linearize_device(pairs, poses, out.data_ptr())
    if not same_stream:
        # the context runs on its own stream (e.g. it was created on the legacy default stream,
        # whose handle is NULL = "private stream"): fall back to a host-side join
        ctx.synchronize()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    res = out.cpu().numpy()
    return res if error_only else res.reshape(n, 91)
