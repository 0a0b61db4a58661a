"""Position sharding and timing reduction for one-process-per-GPU runs (bench.py, N > 1).

The path shards trivially: positions are independent, weights are replicated, there is no
exchange step — so the only distributed operations are a barrier and a MAX-reduce of the timings
(torch.distributed; NCCL on GPUs, gloo in the CPU tests). No collective touches the data path.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of n positions owned by `rank` — the same rule lb2_eval_* uses
    inside one process for several devices (leela_b200/csrc/lb2_api.cu eval_host)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def batch_order(n_sets: int, rank: int) -> list[int]:
    """Weak scaling: every rank evaluates its own batches; start at a different offset so ranks do
    not all evaluate identical positions."""
    return [(rank + i) % n_sets for i in range(n_sets)]


def max_over_ranks(values, dist=None, device=None) -> list[float]:
    """MAX-reduce a list of timings over all ranks (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def aggregate_throughput(units_per_rank_step: int, steps: int, world: int, max_ms: float) -> float:
    """Whole-job units/s = units all ranks processed / the slowest rank's time."""
    return world * units_per_rank_step * steps / (max_ms * 1e-3)


def gather_sharded(local: np.ndarray, n: int, rank: int, world: int, dist=None) -> np.ndarray | None:
    """Reassemble per-rank result slices on rank 0 (used by the sharding tests)."""
    if dist is None or world == 1:
        return local
    import torch
    per = (n + world - 1) // world
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    out = [torch.zeros_like(torch.from_numpy(pad)) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.from_numpy(pad), out, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o.numpy() for o in out])[:n]
