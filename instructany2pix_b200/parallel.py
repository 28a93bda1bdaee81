"""Multi-GPU plumbing: whole sampling trajectories shard by request (prompt x seed); nothing is exchanged inside the loop.

One process per GPU (``torchrun``), ``torch.distributed`` over NCCL (NVLink 5 / NVSwitch) for the single end-of-job
gather of the final latents / images (a few MiB per image).  Seeds derive from the GLOBAL request index so the output is
identical for every GPU count.  The reference has no inference parallelism at all (single process, one
GPU: pipeline.py:124,131); this replaces "run the pipeline N times".
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of global request indices to ranks (load imbalance <= 1 request)."""
    return list(range(rank, n_items, world))


def request_seed(base_seed: int, global_index: int) -> int:
    return base_seed + global_index


def gather_in_order(local: torch.Tensor, local_indices: Sequence[int], n_items: int, dst: int = 0):
    """Gather per-rank results [len(local_indices), ...] to rank ``dst`` as one tensor [n_items, ...] in global order.
    Works with any backend (NCCL on GPUs, gloo in the CPU tests); returns None on other ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        out[list(local_indices)] = local
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_items + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][: len(idx)]
    return out


def run_sharded(n_items: int, batch: int, work: Callable[[List[int]], torch.Tensor]):
    """Run ``work(global_indices) -> Tensor[len, ...]`` over this rank's share in chunks of ``batch``; gather on rank 0."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    mine = shard_indices(n_items, rank, world)
    outs = [work(mine[i:i + batch]) for i in range(0, len(mine), batch)]
    local = torch.cat(outs, 0) if outs else None
    if local is None:   # a rank with no work still has to join the gather with the right trailing shape
        probe = work([])
        local = probe
    return gather_in_order(local, mine, n_items)
