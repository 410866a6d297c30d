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
    """Returns the job-wide statistics vector (the input is left untouched).  ONE collective: every rank's vector is
    gathered (world x 128 bytes) and reduced locally — SUM for the additive slots, MIN / MAX for slots 3 / 4 — instead
    of three all-reduces with different operators (on a GPU whose SMs are all held by a persistent step kernel every
    NCCL kernel waits for a gap between steps, so the count of collectives is what costs).  ``out``: optional
    preallocated ``[(world + 1) * n]`` buffer of the same dtype / device (result in ``out[:n]``), so that the call
    allocates nothing when it is issued on a side stream inside a timed region."""
    n = stats.numel()
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if out is None or out.numel() < (world + 1) * n:
        out = torch.empty((world + 1) * n, dtype=stats.dtype, device=stats.device)
    res = out[:n]
    if world == 1:
        res.copy_(stats)
        return res
    gathered = out[n:(world + 1) * n].view(world, n)
    try:        # one contiguous output: no temporaries, no per-rank copy kernels (NCCL; gloo of recent PyTorch too)
        dist.all_gather_into_tensor(out[n:(world + 1) * n], stats.contiguous(), group=group)
    except (RuntimeError, NotImplementedError, AttributeError):
        dist.all_gather(list(gathered.unbind(0)), stats.contiguous(), group=group)
    torch.sum(gathered, dim=0, out=res)
    res[STATS_MIN:STATS_MIN + 1].copy_(gathered[:, STATS_MIN].min().reshape(1))
    res[STATS_MAX:STATS_MAX + 1].copy_(gathered[:, STATS_MAX].max().reshape(1))
    return res


def summarize(stats: torch.Tensor) -> dict:
    s = stats.detach().cpu().double()
    n = max(float(s[0]), 1.0)
    mean = float(s[1]) / n
    var = max(float(s[2]) / n - mean * mean, 0.0)
    return {"count": float(s[0]), "mean": mean, "std": var ** 0.5, "min": float(s[3]), "max": float(s[4]),
            "episodes_done": float(s[5]), "n_projected": float(s[6]), "n_maxiter": float(s[7]),
            "components_sum": [float(s[8]), float(s[9]), float(s[10])]}
