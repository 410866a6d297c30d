"""Multi-agent EVChargingEnv (one agent per EVSE) on the batched engine.

Mirrors ``sustaingym/envs/evcharging/multiagent_env.py``: PettingZoo ``ParallelEnv`` API, every agent
gets the flattened aggregate observation; with ``periods_delay > 0`` an agent sees the other
stations' ``est_departures``/``demands`` as they were ``periods_delay`` steps ago and its own as
they are now (:102-148); the reward is the global reward divided by the number of agents (:184).

``delay_semantics``.  The reference does NOT deliver that delay: ``EVChargingEnv._get_observation`` returns the same
dict of arrays every step and mutates them in place (env.py:157-163, 383-394); ``MultiAgentEVChargingEnv`` stores
that dict object in its deque (multiagent_env.py:122-123, 130-131), so the "delayed" entries it reads back alias the
current arrays and every agent receives the CURRENT aggregate observation whatever ``periods_delay`` is.
``delay_semantics="reference"`` (default) reproduces that observable behaviour; ``"documented"`` implements what the
docstring describes (others' ``est_departures`` / ``demands`` delayed, own entries, MOER fields and timestep current).

``BatchedMultiAgentEVChargingEnv`` keeps the delay ring buffer on the device and materialises the
``[num_envs, 54, 146]`` per-agent tensor with ``sg_ev_agent_obs`` only when asked;
``MultiAgentEVChargingEnv`` is the reference's single-env dict-of-agents interface.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from ... import _lib, spaces
from .env import BatchedEVChargingEnv
from .event_generation import AbstractTraceGenerator


class BatchedMultiAgentEVChargingEnv:
    def __init__(self, data_generator: AbstractTraceGenerator, num_envs: int = 1, periods_delay: int = 0,
                 moer_forecast_steps: int = 36, project_action_in_env: bool = True, discrete_bins: int = -1,
                 verbose: int = 0, device=None, delay_semantics: str = "reference"):
        if delay_semantics not in ("reference", "documented"):
            raise ValueError("delay_semantics must be 'reference' (aliasing: no effective delay) or 'documented'")
        self.env = BatchedEVChargingEnv(data_generator, num_envs, moer_forecast_steps, project_action_in_env,
                                        verbose, device, auto_reset=False, discrete_bins=discrete_bins)
        self.periods_delay = periods_delay
        self.delay_semantics = delay_semantics
        self.num_envs = num_envs
        self.agents = list(self.env.cn.station_ids)
        self.possible_agents = self.agents
        self.num_agents = len(self.agents)
        n, k = self.env.num_stations, moer_forecast_steps
        cap = data_generator.requested_energy_cap
        low = np.concatenate([np.zeros(n), np.full(n, -288.), np.zeros(k + 2)]).astype(np.float32)
        high = np.concatenate([np.full(n, cap), np.full(n, 288.), np.ones(k + 2)]).astype(np.float32)
        self.flat_observation_space = spaces.Box(low, high, dtype=np.float32)     # flatten_space, sorted keys
        self.agent_action_space = spaces.Discrete(discrete_bins) if discrete_bins > 0 else spaces.Box(0., 1., shape=(1,))
        self._ring: list[torch.Tensor] = []
        self._agent_obs = None

    def _push(self, obs: torch.Tensor, init: bool) -> torch.Tensor:
        """Returns the aggregate observation the *other* agents see (delayed), updating the ring."""
        if self.periods_delay == 0 or self.delay_semantics == "reference":
            return obs
        if init:
            self._ring = [obs.clone() for _ in range(self.periods_delay)]      # multiagent_env.py:119-123
            return obs
        past = self._ring.pop(0)
        self._ring.append(obs.clone())
        return past

    def reset(self, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        self._past = self._push(obs, init=True)
        return obs, info

    def step(self, actions):
        """``actions``: [B, 54] (each agent's scalar action in its column). Returns the aggregate obs,
        per-agent reward (= global / 54), terminated, truncated, info."""
        obs, rew, term, trunc, info = self.env.step(actions)
        if isinstance(obs, np.ndarray):
            obs_t = self.env.obs
        else:
            obs_t = obs
        self._past = self._push(obs_t, init=False)
        return obs, rew / self.num_agents, term, trunc, info

    def agent_observations(self) -> torch.Tensor:
        """[B, 54, 146] per-agent observations for the current step."""
        e = self.env
        B, n, D = e.num_envs, e.num_stations, e.obs_dim
        if self._agent_obs is None:
            self._agent_obs = torch.empty(B, n, D, dtype=torch.float32, device=e.device)
        with torch.cuda.device(e.device):
            _lib.check(e.lib.sg_ev_agent_obs(e.obs.data_ptr(), self._past.contiguous().data_ptr(), B, n, D,
                                             self._agent_obs.data_ptr(), e._stream()), "sg_ev_agent_obs")
        return self._agent_obs


class MultiAgentEVChargingEnv:
    """Reference-compatible PettingZoo ParallelEnv (multiagent_env.py:18-218) on a 1-env batch."""

    metadata: dict[str, Any] = {}

    def __init__(self, data_generator: AbstractTraceGenerator, periods_delay: int = 0, moer_forecast_steps: int = 36,
                 project_action_in_env: bool = True, discrete_bins: int = -1, verbose: int = 0, device=None,
                 delay_semantics: str = "reference"):
        self._b = BatchedMultiAgentEVChargingEnv(data_generator, 1, periods_delay, moer_forecast_steps,
                                                 project_action_in_env, discrete_bins, verbose, device, delay_semantics)
        self.periods_delay = periods_delay
        self.delay_semantics = delay_semantics
        self.agents = list(self._b.agents)
        self.possible_agents = list(self._b.agents)
        self.observation_spaces = {a: self._b.flat_observation_space for a in self.agents}
        self.action_spaces = {a: self._b.agent_action_space for a in self.agents}
        self._discrete = discrete_bins > 0
        self._seed0, self._episodes = 0, 0

    @property
    def num_agents(self) -> int:
        return len(self.agents)

    def _obss(self) -> dict:
        if self.periods_delay == 0 or self.delay_semantics == "reference":
            flat = self._b.env.obs.cpu().numpy()[0]
            return {a: flat for a in self.agents}
        per = self._b.agent_observations().cpu().numpy()[0]
        return {a: per[i] for i, a in enumerate(self.agents)}

    def _info(self) -> dict:
        e = self._b.env
        bd = e.breakdown[0].cpu().numpy()
        return {"max_profit": float(e._max_profit_per_day[int(e.day[0])]),
                "reward_breakdown": {"profit": float(bd[0]), "carbon_cost": float(bd[1]), "excess_charge": float(bd[2])}}

    def reset(self, seed: int | None = None, options: dict | None = None):
        if seed is None:
            seed_eff = self._seed0 + self._episodes
        else:
            self._seed0, self._episodes, seed_eff = int(seed), 0, int(seed)
        self._episodes += 1
        self._b.reset(seed=seed_eff, options=options)
        self.agents = self.possible_agents[:]
        info = self._info()
        return self._obss(), {a: info for a in self.agents}

    def step(self, actions: dict[str, np.ndarray]):
        n = len(self.possible_agents)
        a = np.zeros((1, n), np.int64 if self._discrete else np.float32)
        for i, agent in enumerate(self.agents):
            a[0, i] = np.asarray(actions[agent]).reshape(-1)[0]
        _, rew, term, trunc, _ = self._b.step(a)
        obss, info = self._obss(), self._info()
        terminated, truncated = bool(term[0]), False
        rewards = {ag: float(rew[0]) for ag in self.agents}
        terms = {ag: terminated for ag in self.agents}
        truncs = {ag: truncated for ag in self.agents}
        infos = {ag: info for ag in self.agents}
        if terminated or truncated:
            self.agents = []
        return obss, rewards, terms, truncs, infos

    def observation_space(self, agent: str):
        return self.observation_spaces[agent]

    def action_space(self, agent: str):
        return self.action_spaces[agent]

    def render(self) -> None:
        pass

    def close(self) -> None:
        pass
