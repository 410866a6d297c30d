"""Evaluation harness for controllers on the engine's envs.

``BaseAlgorithm`` keeps the reference's interface (sustaingym/algorithms/base.py:15-97): subclasses implement
``get_action(observation)``; ``run(seeds)`` plays one episode per seed on a single env and returns a DataFrame
with ``seed``, ``return`` and the final ``info`` entries.  ``run_batched`` is what the engine adds: ALL seeds
are played in lock-step on one batched env (one env instance per seed, one kernel launch per step for all of
them), the policy being a function of the batched CUDA observation tensor — same table, len(seeds) times
fewer launches.
"""
from __future__ import annotations

from collections.abc import Sequence
from copy import deepcopy
from typing import Any, Callable

import numpy as np


class BaseAlgorithm:
    """A controller bound to one env.  Subclasses provide ``get_action``; ``reset`` is called between episodes."""

    def __init__(self, env, multiagent: bool = False):
        self.env = env
        self.multiagent = multiagent

    def get_action(self, observation: dict[str, Any]):
        raise NotImplementedError

    def reset(self) -> None:
        return None

    def _episode(self, seed: int) -> tuple[float, dict]:
        """Plays the episode ``seed`` selects to its end: (sum of rewards, last info)."""
        obs, info = self.env.reset(seed=seed)
        self.reset()
        total, finished = 0.0, False
        while not finished:
            obs, rew, term, trunc, info = self.env.step(self.get_action(obs))
            if self.multiagent:          # PettingZoo dicts: team reward, any agent ending ends the episode
                total += sum(rew.values())
                finished = any(term.values()) or any(trunc.values())
            else:
                total += rew
                finished = bool(term or trunc)
        if self.multiagent and info:
            info = next(iter(info.values()))   # all agents carry the same info
        return total, info

    def run(self, seeds: Sequence[int] | int):
        """One row per seed: ``seed``, ``return`` and the entries of the episode's final ``info``
        (the table of sustaingym/algorithms/base.py:35-97)."""
        import pandas as pd
        rows = []
        for seed in (range(seeds) if isinstance(seeds, int) else seeds):
            total, info = self._episode(int(seed))
            rows.append({"seed": int(seed), "return": total, **{k: deepcopy(v) for k, v in info.items()}})
        return pd.DataFrame(rows, columns=list(rows[0]) if rows else ["seed", "return"])


def run_batched(batched_env, policy: Callable, seeds: Sequence[int] | int, max_steps: int | None = None):
    """Plays ``len(seeds)`` episodes in lock-step on ``batched_env`` (``num_envs == len(seeds)``, auto-reset off).

    ``policy(obs_dict) -> actions``: ``obs_dict`` maps the observation keys to ``[num_envs, ...]`` CUDA views
    (``batched_env.obs_dict()``), the result is a ``[num_envs, n]`` CUDA tensor of the env's action dtype.
    Returns the DataFrame ``BaseAlgorithm.run`` would produce for the same seeds (``return`` accumulated in f64 on
    the device; info columns from ``batched_env.episode_info()``)."""
    import pandas as pd
    import torch
    if isinstance(seeds, int):
        seeds = list(range(seeds))
    seeds = [int(s) for s in seeds]
    if batched_env.num_envs != len(seeds):
        raise ValueError(f"batched env has {batched_env.num_envs} envs for {len(seeds)} seeds")
    if getattr(batched_env, "auto_reset", False):
        raise ValueError("run_batched needs an env built with auto_reset=False (episodes must end)")
    obs, _ = batched_env.reset(seed=np.asarray(seeds))
    ret = torch.zeros(len(seeds), dtype=torch.float64, device=obs.device)
    alive = torch.ones(len(seeds), dtype=torch.bool, device=obs.device)
    steps = 0
    while bool(alive.any()) and (max_steps is None or steps < max_steps):
        actions = policy(batched_env.obs_dict(obs))
        obs, reward, term, trunc, _ = batched_env.step(actions)
        ret += torch.where(alive, reward.to(torch.float64), torch.zeros_like(ret))
        alive &= ~(term | trunc)
        steps += 1
    out: dict[str, list] = {"seed": seeds, "return": ret.cpu().tolist()}
    for key, values in batched_env.episode_info().items():
        out[key] = values
    return pd.DataFrame(out)
