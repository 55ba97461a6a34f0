"""Multi-GPU plumbing of the view-sharded optimisation loop.

One process per GPU (torchrun), a full replica of the Gaussian cloud per GPU, training views
sharded round-robin.  The data path has exactly one exchange per optimizer step: a SUM all-reduce
of the flat per-Gaussian gradient buffer (+ the tiny [n_views,7] pose-gradient table, whose rows are
disjoint across ranks); the x 1/G is folded into the Adam kernel.  Works with any torch.distributed
backend (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import List, Sequence

import torch


def shard_views(n_views: int, world_size: int, rank: int) -> List[int]:
    """view v -> rank v mod G."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return [v for v in range(n_views) if v % world_size == rank]


def view_for_step(n_views: int, world_size: int, rank: int, step: int) -> int:
    """The view rank `rank` renders at optimizer step `step` (cycles through its shard; ranks with
    an empty shard fall back to the global round-robin so every rank always has work)."""
    mine = shard_views(n_views, world_size, rank)
    if not mine:
        return (step * world_size + rank) % n_views
    return mine[step % len(mine)]


def allreduce_sum_(tensors: Sequence[torch.Tensor], group=None) -> None:
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def init_from_env(backend: str = "nccl"):
    """torchrun-style initialisation; returns (rank, local_rank, world_size)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world
