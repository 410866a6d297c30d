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


def all_reduce_stats(stats: torch.Tensor, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Returns the job-wide statistics vector (the input is left untouched).  ``out``: optional preallocated
    ``[SG_STATS_LEN + 2]`` buffer of the same dtype/device (result in ``out[:SG_STATS_LEN]``, the last two
    entries are the MIN / MAX scratch), so that the call allocates nothing — needed when it is issued on a side
    stream inside a timed region (a caching-allocator miss on a new stream costs a cudaMalloc)."""
    n = stats.numel()
    if out is None:
        out = torch.empty(n + 2, dtype=stats.dtype, device=stats.device)
    res, mn, mx = out[:n], out[n:n + 1], out[n + 1:n + 2]
    res.copy_(stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return res
    mn.copy_(stats[STATS_MIN:STATS_MIN + 1])
    mx.copy_(stats[STATS_MAX:STATS_MAX + 1])
    dist.all_reduce(res, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    res[STATS_MIN:STATS_MIN + 1].copy_(mn)
    res[STATS_MAX:STATS_MAX + 1].copy_(mx)
    return res


def summarize(stats: torch.Tensor) -> dict:
    s = stats.detach().cpu().double()
    n = max(float(s[0]), 1.0)
    mean = float(s[1]) / n
    var = max(float(s[2]) / n - mean * mean, 0.0)
    return {"count": float(s[0]), "mean": mean, "std": var ** 0.5, "min": float(s[3]), "max": float(s[4]),
            "episodes_done": float(s[5]), "n_projected": float(s[6]), "n_maxiter": float(s[7]),
            "components_sum": [float(s[8]), float(s[9]), float(s[10])]}
