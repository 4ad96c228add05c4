"""Multi-GPU plumbing: streams are independent, so ranks take contiguous stream ranges and never exchange
data on the data path. torch.distributed is only used for the barrier and for reducing the
timing / unit counters that bench.py reports (max over ranks, sum of units)."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of `n` streams for `rank` of `world`."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stream_ids(per_rank: int, rank: int):
    """Weak scaling: every rank synthesises its own `per_rank` stream ids."""
    return list(range(rank * per_rank, (rank + 1) * per_rank))


def _reduce(value: float, op_name: str) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def all_max(value: float) -> float:
    return _reduce(value, "MAX")


def all_sum(value: float) -> float:
    return _reduce(value, "SUM")
