"""Multi-GPU plumbing: agents shard by contiguous slice, one process per GPU, no per-tick communication.

Agents never see each other (they only read static track geometry), so the tick needs no collective.  The
only exchange the population learners need is per generation: every rank must see every candidate's scalar
fitness to rank the population (CMA-ES `tell`, CmaEsSolverTorch.cpp:81-128; genetic top-5, Mating.hpp:108-166).
That is one all-gather of f32[population] over NCCL (NVLink 5 / NVSwitch) -- 32 MB at 8M agents.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, rank: int, world: int):
    """contiguous slice [lo, hi) of the global agent range owned by `rank` (sizes differ by at most 1)"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_fitness(local: torch.Tensor, n_total: int | None = None) -> torch.Tensor:
    """Global fitness vector f32[n_total], identical on every rank, ordered by global agent id.
    `local` is this rank's slice (sizes may differ by one between ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.clone()
    world, rank = dist.get_world_size(), dist.get_rank()
    if n_total is None:
        t = torch.tensor([local.numel()], device=local.device, dtype=torch.int64)
        dist.all_reduce(t)
        n_total = int(t.item())
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == width for lo, hi in sizes):
        out = torch.empty(n_total, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = torch.zeros(width, device=local.device, dtype=local.dtype)
    pad[: local.numel()] = local
    buf = torch.empty(world * width, device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(buf, pad)
    return torch.cat([buf[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def global_ranking(local_fitness: torch.Tensor, n_total: int | None = None, descending: bool = True):
    """(fitness, order): the gathered fitness and the indices that sort it.  Ties are broken by global agent
    id (stable sort) so every rank derives the same parents."""
    f = all_gather_fitness(local_fitness, n_total)
    order = torch.sort(f, descending=descending, stable=True).indices
    return f, order


def top_k(local_fitness: torch.Tensor, k: int, n_total: int | None = None):
    """global top-k (value, global agent id), e.g. the genetic learner's 5 parents (Mating.hpp:118-119)"""
    f, order = global_ranking(local_fitness, n_total)
    idx = order[:k]
    return f[idx], idx
