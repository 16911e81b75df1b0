"""Multi-GPU: one process per GPU, independent batch shards, ONE NCCL broadcast of the packed weight
arenas at load and no per-step collective.  Every latent sample is independent
through all DDIM steps, decode and vocode, so ranks never exchange data on the hot path."""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of ``n_items`` units for ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_arena_bcast(device, src: int = 0):
    """Hook for NativeLatentDiffusion(arena_bcast=...): rank ``src`` uploads its packed arena, every
    other rank receives it over NCCL (NVLink/NVSwitch) instead of packing + uploading its own."""
    def bcast(name: str, cpu_arena: torch.Tensor, nbytes: int) -> torch.Tensor:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return cpu_arena.to(device)
        if dist.get_rank() == src:
            t = cpu_arena.to(device)
        else:
            t = torch.empty(nbytes, dtype=torch.uint8, device=device)
        dist.broadcast(t, src=src)
        return t
    return bcast


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
