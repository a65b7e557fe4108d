"""Multi-GPU plumbing: one process per GPU, envs sharded contiguously by *global* env id.

Envs are independent (nothing in ``Tetris.move`` touches another env, ``game/tetris.py:354-422``), so the path
shards with no data-path collective; every RNG stream is keyed by the global env id, so results do not depend on
the number of ranks.  The only collective is one SUM all-reduce of the 8 x int64 episode-statistics vector per
rollout (64 bytes: latency-bound, NCCL over NVLink on GPUs, gloo in the CPU test tier).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(env_base, num_envs) of `rank`: contiguous blocks, remainder spread over the first ranks."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    q, r = divmod(int(total_envs), int(world_size))
    base = rank * q + min(rank, r)
    return base, q + (1 if rank < r else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns
    (rank, world_size, local_rank); a single process needs no initialisation."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def all_reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """SUM over ranks of the int64[8] statistics vector (returns a new tensor; identity when not distributed)."""
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def sharded_env(total_envs: int, L: int, M: int, seed: int = 0, config_pool=None, **kw):
    """A ``BatchedTetris`` holding this rank's shard of `total_envs` envs on its local GPU."""
    from .batched import BatchedTetris
    rank, world, local = init_from_env()
    base, count = shard_bounds(total_envs, world, rank)
    return BatchedTetris(count, L, M, device=torch.device("cuda", local), seed=seed, config_pool=config_pool,
                         env_base=base, **kw)
