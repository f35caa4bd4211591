"""Multi-GPU plumbing of the path: sequences are independent units, so the forward shards the
batch across ranks with NO data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) carries
the barrier and the max-over-ranks reduction of the timed region and, for training, the one
real exchange step of the path: the sum-all-reduce of the parameter gradients (`allreduce_gradients`)."""
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


def allreduce_gradients(params, world: "int | None" = None, group=None, average: bool = True) -> int:
    """Data-parallel gradient exchange (the reference's `nn.DataParallel` reduce, train.py:205 + :258): every
    `p.grad` is packed into ONE flat fp32 bucket (170 MB for DSTformer-base), summed over the ranks with a single
    all-reduce (NCCL over NVLink/NVSwitch on GPUs) and written back, divided by the world size when `average`.
    Returns the number of elements exchanged.  A no-op outside an initialised process group."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    n_ranks = world if world is not None else dist.get_world_size(group)
    if n_ranks == 1:
        return 0
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(n_ranks)
    for g, r in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(r)
    return int(flat.numel())
