"""Multi-GPU plumbing: whole sampling trajectories shard by request (prompt x seed); nothing is exchanged inside the loop.

One process per GPU (``torchrun``), ``torch.distributed`` over NCCL (NVLink 5 / NVSwitch) for the single end-of-job
gather of the final latents / images (a few MiB per image).  Seeds derive from the GLOBAL request index, and requests are dealt
out in whole BATCHES (block-cyclic, block = per-GPU batch): a request shares its kernel launches with the same neighbours, in the
same batch slot, for every GPU count, so the output is bit-identical for 1, 2, 4 or 8 GPUs (the kernels are
bit-reproducible for a fixed shape, not across batch compositions -- DESIGN.md section 5).  The reference has no inference parallelism at all (single process, one
GPU: pipeline.py:124,131); this replaces "run the pipeline N times".
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int, block: int = 1) -> List[int]:
    """Block-cyclic assignment of global request indices to ranks: consecutive blocks of ``block`` requests (one per-GPU batch) go
    round the ranks, so batch k always holds requests [k * block, (k + 1) * block) whatever the GPU count (load imbalance <= one
    batch).  block = 1 is plain round-robin."""
    block = max(int(block), 1)
    return [i for b0 in range(rank * block, n_items, world * block) for i in range(b0, min(b0 + block, n_items))]


def request_seed(base_seed: int, global_index: int) -> int:
    return base_seed + global_index


def gather_in_order(local: torch.Tensor, local_indices: Sequence[int], n_items: int, dst: int = 0, block: int = 1):
    """Gather per-rank results [len(local_indices), ...] to rank ``dst`` as one tensor [n_items, ...] in global order.
    Works with any backend (NCCL on GPUs, gloo in the CPU tests); returns None on other ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        out[list(local_indices)] = local
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    per = max(len(shard_indices(n_items, r, world, block)) for r in range(world))
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world, block)
        out[idx] = bufs[r][: len(idx)]
    return out


def run_sharded(n_items: int, batch: int, work: Callable[[List[int]], torch.Tensor]):
    """Run ``work(global_indices) -> Tensor[len, ...]`` over this rank's share in chunks of ``batch``; gather on rank 0."""
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    mine = shard_indices(n_items, rank, world, batch)
    outs = [work(mine[i:i + batch]) for i in range(0, len(mine), batch)]
    local = torch.cat(outs, 0) if outs else None
    if local is None:   # a rank with no work still has to join the gather with the right trailing shape
        probe = work([])
        local = probe
    return gather_in_order(local, mine, n_items, block=batch)
