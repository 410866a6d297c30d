"""Multi-GPU plumbing: one process per GPU, envs sharded by contiguous id ranges, one collective.

Env instances are independent (each reference env owns its simulator: ``evcharging/env.py:324-329``,
``building/env.py:108-203``, ``cogen/env.py:211-225``), so the batch shards with no data-path
collective.  The only exchange is the aggregate reward-statistics vector the kernels accumulate
(``SG_STATS_LEN`` doubles: count, sum, sum of squares, min, max, episode/solver counters, reward
components): SUM for the additive slots, MIN/MAX for slots 3/4.  Works with NCCL (GPU tensors) and
gloo (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

STATS_MIN, STATS_MAX = 3, 4


def shard_range(num_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [lo, hi) env-id range of ``rank``; the first ``num_envs % world_size`` ranks get one more."""
    base, rem = divmod(num_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Returns the job-wide statistics vector (the input is left untouched)."""
    out = stats.clone()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return out
    mn, mx = out[STATS_MIN:STATS_MIN + 1].clone(), out[STATS_MAX:STATS_MAX + 1].clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out[STATS_MIN], out[STATS_MAX] = mn[0], mx[0]
    return out


def summarize(stats: torch.Tensor) -> dict:
    s = stats.detach().cpu().double()
    n = max(float(s[0]), 1.0)
    mean = float(s[1]) / n
    var = max(float(s[2]) / n - mean * mean, 0.0)
    return {"count": float(s[0]), "mean": mean, "std": var ** 0.5, "min": float(s[3]), "max": float(s[4]),
            "episodes_done": float(s[5]), "n_projected": float(s[6]), "n_maxiter": float(s[7]),
            "components_sum": [float(s[8]), float(s[9]), float(s[10])]}
