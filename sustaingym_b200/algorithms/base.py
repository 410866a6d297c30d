"""Evaluation harness for controllers on the engine's envs.

``BaseAlgorithm`` keeps the reference's interface (sustaingym/algorithms/base.py:15-97): subclasses implement
``get_action(observation)``; ``run(seeds)`` plays one episode per seed on a single env and returns a DataFrame
with ``seed``, ``return`` and the final ``info`` entries.  ``run_batched`` is what the engine adds: ALL seeds
are played in lock-step on one batched env (one env instance per seed, one kernel launch per step for all of
them), the policy being a function of the batched CUDA observation tensor — same table, len(seeds) times
fewer launches.
"""
from __future__ import annotations

from collections import defaultdict
from collections.abc import Sequence
from copy import deepcopy
from typing import Any, Callable

import numpy as np


class BaseAlgorithm:
    def __init__(self, env, multiagent: bool = False):
        self.env = env
        self.multiagent = multiagent

    def get_action(self, observation: dict[str, Any]):
        raise NotImplementedError

    def reset(self) -> None:
        """Resets the algorithm at the end of each episode."""

    def run(self, seeds: Sequence[int] | int):
        """base.py:35-97: one episode per seed; DataFrame with ``seed``, ``return`` and the final info entries."""
        import pandas as pd
        if isinstance(seeds, int):
            seeds = list(range(seeds))
        results: dict[str, list] = defaultdict(list)
        for seed in seeds:
            results["seed"].append(seed)
            ep_return = 0.0
            obs, _ = self.env.reset(seed=seed)
            self.reset()
            done = False
            info: dict = {}
            while not done:
                action = self.get_action(obs)
                obs, reward, terminated, truncated, info = self.env.step(action)
                assert isinstance(reward, dict) == self.multiagent
                if self.multiagent:
                    reward = sum(reward.values())
                    done = any(terminated.values()) or any(truncated.values())
                else:
                    done = terminated or truncated
                ep_return += reward
            results["return"].append(ep_return)
            if self.multiagent:
                info = info[list(info.keys())[0]]
            for key, value in info.items():
                results[key].append(deepcopy(value))
        return pd.DataFrame(results)


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
