"""Baseline controllers for EVChargingEnv (sustaingym/algorithms/evcharging/baselines.py:22-50), for the
single-env interface and, as device policies, for ``run_batched``.  (The reference's MPC / offline-optimal
controllers are cvxpy programs over the same network constraints; they are not part of the engine.)"""
from __future__ import annotations

from typing import Any

import numpy as np

from ..base import BaseAlgorithm

MAX_ACTION = 1      # continuous action space
D_MAX_ACTION = 4    # DiscreteActionWrapper


class GreedyAlgorithm(BaseAlgorithm):
    """Full pilot wherever the demand is non-zero (baselines.py:22-35)."""

    def __init__(self, env):
        super().__init__(env, multiagent=False)
        self.continuous_action_space = hasattr(env.action_space, "low")
        self.max_action = MAX_ACTION if self.continuous_action_space else D_MAX_ACTION

    def get_action(self, observation: dict[str, Any]) -> np.ndarray:
        return np.where(observation["demands"] > 0, self.max_action, 0)


class RandomAlgorithm(BaseAlgorithm):
    """Uniformly random action (baselines.py:38-50)."""

    def __init__(self, env, seed: int | None = None):
        super().__init__(env, multiagent=False)
        self.continuous_action_space = hasattr(env.action_space, "low")
        self.rng = np.random.default_rng(seed)

    def get_action(self, observation: dict[str, Any]) -> np.ndarray:
        n = self.env.num_stations
        if self.continuous_action_space:
            return self.rng.random(size=n).astype(np.float32)
        return self.rng.choice(D_MAX_ACTION + 1, size=n).astype(np.float32)


def greedy_policy(batched_env):
    """Device policy for ``run_batched``: the greedy rule on the batched observation."""
    import torch
    discrete = batched_env.discrete_bins > 0
    hi = D_MAX_ACTION if discrete else MAX_ACTION
    dtype = torch.int64 if discrete else torch.float32

    def policy(obs: dict):
        return (obs["demands"] > 0).to(dtype) * hi
    return policy
