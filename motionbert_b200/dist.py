"""Multi-GPU plumbing of the path: sequences are independent units, so the forward shards the
batch across ranks with NO data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) only carries
the barrier and the max-over-ranks reduction of the timed region."""
from __future__ import annotations

import os

import torch


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n_global: int, rank: int, world: int) -> "tuple[int, int]":
    """Contiguous, balanced [lo, hi) slice of n_global independent sequences owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init(backend: str, device: "torch.device | None" = None):
    import torch.distributed as dist
    if dist.is_initialized():
        return dist
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend, **kw)
    return dist


def max_over_ranks(value: float, device: "torch.device | str" = "cpu") -> float:
    """Time of the slowest rank (multi-GPU numbers are the max over ranks, never a wall clock)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: "torch.device | str" = "cpu") -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
