"""Multi-GPU plumbing: the path shards by env / record index with NO data-path collective.

One process per GPU (torchrun); rank r of G owns the contiguous block [r*N/G, (r+1)*N/G) of cars or tub records.
The only collectives are (i) an all-reduce of the per-step statistics vector (16 x int64 = 128 B per rank)
and (ii), in tests, a gather of shard outputs to check them against the single-process result.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block of global indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank statistics vector over all ranks (the system's only hot-loop collective: 128 bytes)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        stats = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value: float, device=None) -> float:
    """Max of a python float over ranks (used for the timing rule: a multi-GPU step costs its slowest rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def gather_shards(local: torch.Tensor, n_total: int) -> torch.Tensor | None:
    """Test helper: concatenate contiguous shards on rank 0 (rows may differ by one between ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_rows = max(e - s for s, e in sizes)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != 0:
        return None
    return torch.cat([b[: e - s] for b, (s, e) in zip(bufs, sizes)], dim=0)
