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


# ----------------------------------------------------------------------------------------------
# peer-visible device buffers (CUDA IPC) for the fused reduce-scatter -> Adam -> all-gather kernel
# ----------------------------------------------------------------------------------------------
class _DevMem:
    """Exposes a raw device allocation through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False),
                                         "version": 2}


class PeerBuffer:
    """A flat fp32 buffer allocated by libgsb200.so (cudaMalloc) on this rank's GPU and opened, through
    CUDA IPC handles exchanged over torch.distributed, by every other rank of the node."""

    def __init__(self, n_floats: int, device, group=None):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        self.L, self.n = L, n_floats
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        with torch.cuda.device(device):
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.check(L.gsb_ipc_alloc(n_floats * 4, ctypes.byref(ptr), handle), "gsb_ipc_alloc")
            self.local_ptr = ptr.value
            handles = [None] * world
            dist.all_gather_object(handles, handle.raw, group=group)
            self.ptrs = []
            for r in range(world):
                if r == rank:
                    self.ptrs.append(self.local_ptr)
                else:
                    p = ctypes.c_void_p()
                    _lib.check(L.gsb_ipc_open(handles[r], ctypes.byref(p)), f"gsb_ipc_open(rank {r})")
                    self.ptrs.append(p.value)
        self._mem = _DevMem(self.local_ptr, n_floats)
        self.tensor = torch.as_tensor(self._mem, device=device)
        self.rank, self.world = rank, world

    def ptr_array(self):
        import ctypes
        return (ctypes.c_void_p * self.world)(*self.ptrs)

    def close(self, group=None):
        """Collective: drop the torch view, close the peer mappings, then (after a barrier, so that nobody still maps
        it) free the local allocation.  The buffer must not be used afterwards."""
        import torch.distributed as dist
        if self.local_ptr is None:
            return
        torch.cuda.synchronize()
        self.tensor = None
        self._mem = None
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                self.L.gsb_ipc_close(p)
        if dist.is_initialized():
            dist.barrier(group=group)
        self.L.gsb_ipc_free(self.local_ptr)
        self.local_ptr, self.ptrs = None, []


def shard_bounds(total: int, world: int, rank: int, align: int = 768):
    """Contiguous shard [lo, hi) of a flat buffer of `total` floats; boundaries are multiples of `align`
    (a multiple of every row length with a per-row multiplier and of the 128-bit vector width)."""
    per = (total + world - 1) // world
    per = (per + align - 1) // align * align
    lo = min(total, rank * per)
    hi = min(total, lo + per)
    return lo, hi
