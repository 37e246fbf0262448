"""Dataset-level channel / scalar statistics across ranks (BASELINE config 3: "NCCL stats all-reduce").

Per-segment normalisation in the reference is local to each segment (process.py:36-38,47,55,60,65,70,76), so these
statistics never change the `.npz` contents; they are an extra summary a trainer may use.  The accumulator is a
[(9 + S), 5] float64 tensor = {count, sum, sum of squares, min, max}; the exchange is one SUM all-reduce over the first
three columns plus one MIN and one MAX over the last two (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def allreduce_stats(st, dist=None, group=None):
    """In-place reduction of a stats tensor across the default (or given) process group; returns `st`."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return st
    head = st[:, 0:3].contiguous()
    mn = st[:, 3].contiguous()
    mx = st[:, 4].contiguous()
    dist.all_reduce(head, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    st[:, 0:3] = head
    st[:, 3] = mn
    st[:, 4] = mx
    return st


def finalize_stats(st):
    """-> dict(count, mean, std, min, max) as float64 tensors / arrays of length 9 + S."""
    cnt = st[:, 0].clip(1.0) if hasattr(st, "clip") else st[:, 0]
    mean = st[:, 1] / cnt
    var = (st[:, 2] / cnt - mean * mean)
    var = var.clip(0.0) if hasattr(var, "clip") else var
    return {"count": st[:, 0], "mean": mean, "std": var ** 0.5, "min": st[:, 3], "max": st[:, 4]}


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous batch shard of rank r: [r*N/R, (r+1)*N/R)."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi
