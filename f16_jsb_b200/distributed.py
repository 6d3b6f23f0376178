"""Multi-GPU plumbing: one process per GPU, contiguous env shards, no data-path collective.

Envs are independent, so each rank owns envs [rank*N/W, (rank+1)*N/W) with its own
state/obs buffers; Philox keys use the global env id, so results do not depend on the world size.
The only collective is an all-reduce (sum) of the 8-double rollout-statistics vector once per log
interval - NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of global env ids owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's env vars; returns (rank, local_rank, world_size)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank statistics vector over all ranks (in place); identity for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value: float, device=None) -> float:
    """Max of a scalar over ranks (timing: a multi-GPU step takes as long as its slowest rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
