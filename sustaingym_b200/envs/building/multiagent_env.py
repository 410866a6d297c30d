"""Multi-agent BuildingEnv (one agent per zone with an AC unit) — mirrors
``sustaingym/envs/building/multiagent_env.py`` on top of the batched engine."""
from __future__ import annotations

from typing import Any, Mapping

import numpy as np

from ... import spaces
from .env import BuildingEnv


class MultiAgentBuildingEnv:
    def __init__(self, parameters: dict[str, Any], device=None, positional_keys: bool = False) -> None:
        """``positional_keys``: the reference reads ``actions[i]`` with i the ENUMERATION index of the agent
        (multiagent_env.py:84-85), which equals the agent id only when every zone has an AC unit.  Default here:
        the PettingZoo convention (the dict is keyed by agent id, as ``parallel_api_test`` drives it);
        ``positional_keys=True`` reproduces the reference's indexing."""
        self.positional_keys = positional_keys
        self.single_env = BuildingEnv(parameters, device)
        self.possible_agents = np.nonzero(self.single_env.ac_map)[0].tolist()      # multiagent_env.py:53
        self.agents = self.possible_agents[:]
        self.observation_spaces = {a: self.single_env.observation_space for a in self.agents}
        if self.single_env.is_continuous_action:
            self.action_spaces = {a: spaces.Box(-1.0, 1.0, shape=(1,), dtype=np.float32) for a in self.agents}
        else:
            self.action_spaces = {a: self.single_env.action_space[a] for a in self.agents}

    @property
    def num_agents(self) -> int:
        return len(self.agents)

    def step(self, actions: Mapping[int, np.ndarray]):
        dt = np.float32 if self.single_env.is_continuous_action else np.int64
        action = np.zeros(self.single_env.n, dtype=dt)
        for i, agent in enumerate(self.agents):
            v = actions[i] if self.positional_keys else actions[agent]
            action[agent] = np.asarray(v).reshape(-1)[0]
        obs, reward, terminated, truncated, info = self.single_env.step(action)
        self._state = obs
        obss = {a: obs for a in self.agents}
        rewards = {a: reward for a in self.agents}
        terms = {a: terminated for a in self.agents}
        truncs = {a: truncated for a in self.agents}
        infos = {a: info for a in self.agents}
        if terminated or truncated:
            self.agents = []
        return obss, rewards, terms, truncs, infos

    def reset(self, seed: int | None = None, options: dict | None = None):
        obs, info = self.single_env.reset(seed=seed, options=options)
        self._state = obs
        self.agents = self.possible_agents[:]
        return {a: obs for a in self.agents}, {a: info for a in self.agents}

    def state(self) -> np.ndarray:
        return self._state

    def observation_space(self, agent: int):
        return self.observation_spaces[agent]

    def action_space(self, agent: int):
        return self.action_spaces[agent]

    def render(self) -> None:
        pass

    def close(self) -> None:
        self.single_env.close()
