"""Dataset-level channel / scalar statistics across ranks (BASELINE config 3: "NCCL stats all-reduce").

Per-segment normalisation in the reference is local to each segment (process.py:36-38,47,55,60,65,70,76), so these
statistics never change the `.npz` contents; they are an extra summary a trainer may use.  The accumulator is a
[(9 + S), 5] float64 tensor = {count, sum, sum of squares, min, max}.  The exchange is ONE collective: an all-gather of
the 1.8 KB accumulators (NCCL on GPUs, gloo in the CPU tests), after which every rank reduces the R copies locally --
SUM over the first three columns, MIN and MAX over the last two, in rank order, so every rank holds bit-identical results.
"""
from __future__ import annotations


def allreduce_stats(st, dist=None, group=None):
    """In-place reduction of a stats tensor across the default (or given) process group; returns `st`."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return st
    world = dist.get_world_size(group)
    src = st.contiguous()
    parts = [src.new_empty(src.shape) for _ in range(world)]
    dist.all_gather(parts, src, group=group)
    allr = parts[0].new_empty((world,) + tuple(src.shape))
    for r, p in enumerate(parts):
        allr[r] = p
    st[:, 0:3] = allr[:, :, 0:3].sum(dim=0)
    st[:, 3] = allr[:, :, 3].min(dim=0).values
    st[:, 4] = allr[:, :, 4].max(dim=0).values
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
