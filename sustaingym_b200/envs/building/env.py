"""Batched BuildingEnv on the B200 engine.

``BatchedBuildingEnv(parameters, num_envs)`` takes the same ``parameters`` dict the reference's
``BuildingEnv`` takes (``sustaingym/envs/building/env.py:110-203``; produced by the reference's
``ParameterGenerator`` or by its restatement here, ``envs/building/parameters.py`` — or just the name of one of the 16
prototype buildings, e.g. ``BuildingEnv("OfficeSmall")``) and advances ``num_envs`` independent copies per kernel launch.
``BuildingEnv`` is the ``num_envs = 1`` Gymnasium view.  All per-step arithmetic runs in
``sg_building_step`` / ``sg_building_rollout`` / ``sg_building_reset``.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from ... import _lib, spaces
from .._state import StateDictMixin, nvtx_method

OCCU_COEF_LINEAR = 7.139322   # env.py:99
DISCRETE_LENGTH = 100         # env.py:102
SCALING_FACTOR = 24           # env.py:105


def discretize(A: np.ndarray, B: np.ndarray, D: np.ndarray, dt: float):
    """Zero-order-hold discretisation done once at construction (env.py:196-203):
    ``A_d = expm(A dt)``, ``BD_d = A^-1 (A_d - I) [D | B]``."""
    from scipy.linalg import expm
    A = np.asarray(A, np.float64)
    BD = np.hstack((np.asarray(D, np.float64)[:, None], np.asarray(B, np.float64)))
    A_d = expm(A * dt)
    BD_d = np.linalg.inv(A) @ (A_d - np.eye(A.shape[0])) @ BD
    return A_d, BD_d


def seed_to_epoch(seed: int, episode_len: int, time_res: int, length_of_weather: int) -> int:
    """Weather cursor of ``reset(seed=...)`` (env.py:343-345)."""
    num_days_normalizer = ((episode_len * time_res) // 86_400) * 365
    return min(int((seed / num_days_normalizer) * length_of_weather), length_of_weather - 1)


class BatchedBuildingEnv(StateDictMixin):

    _STATE_TENSORS = ('state', 'epoch', 'runs', 'reward', 'done', 'stats')          # checkpointed by state_dict() (envs/_state.py)
    def __init__(self, parameters: dict[str, Any], num_envs: int = 1, device=None, auto_reset: bool = False,
                 collect_stats: bool = False, tensor_cores: bool = True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedBuildingEnv needs a CUDA device (no CPU fallback)")
        self.device = dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if isinstance(parameters, str):     # a prototype-building name: the reference's test set-up (tests/test_building.py:17-19)
            from .parameters import ParameterGenerator
            parameters = ParameterGenerator(parameters, "Hot_Dry", "Tucson")
        p = self.parameters = parameters
        self.num_envs = B = int(num_envs)
        self.n = n = int(p["n"])
        self.target = np.asarray(p["target"], np.float64)
        self.ac_map = np.asarray(p["ac_map"], np.float64)
        self.maxpower = p["max_power"]
        self.temp_range = p["temp_range"] if "temp_range" in p else (-40, 40)
        self.reward_pnorm = float(p["reward_pnorm"])
        self.is_continuous_action = bool(p["is_continuous_action"])
        self.timestep = int(p["time_resolution"])
        self.episode_len = int(p["episode_len"])
        self.length_of_weather = T = len(p["out_temp"])
        self.q_rate = (1 - p["reward_beta"]) * SCALING_FACTOR
        self.error_rate = p["reward_beta"]
        self.auto_reset = auto_reset
        if "A_d" in p and "BD_d" in p:
            self.A_d, self.BD_d = np.asarray(p["A_d"], np.float64), np.asarray(p["BD_d"], np.float64)
        else:
            self.A_d, self.BD_d = discretize(p["A"], p["B"], p["D"], self.timestep)

        # ---- spaces (env.py:143-176); the bounds quirk of the reference is kept as is -----------
        self.Qlow = -self.ac_map.astype(np.float32)
        self.Qhigh = self.ac_map.astype(np.float32)
        if self.is_continuous_action:
            self.single_action_space = spaces.Box(self.Qlow, self.Qhigh, dtype=np.float32)
        else:
            self.single_action_space = spaces.MultiDiscrete(
                (self.Qhigh * DISCRETE_LENGTH - self.Qlow * DISCRETE_LENGTH).astype(np.int64))
        min_T, max_T = self.temp_range
        heat_max = 1000
        self.low = np.concatenate([np.ones(n + 1) * min_T, [0], [min_T],
                                   [-min_T * OCCU_COEF_LINEAR / 1000]]).astype(np.float32)
        self.high = np.concatenate([np.ones(n + 1) * max_T, [heat_max], [max_T], [heat_max]]).astype(np.float32)
        self.single_observation_space = spaces.Box(self.low, self.high, dtype=np.float32)

        # ---- device constants -----------------------------------------------------------------------
        weather = np.stack([np.asarray(p["out_temp"], np.float64), np.asarray(p["ground_temp"], np.float64),
                            np.asarray(p["ghi"], np.float64), np.asarray(p["metabolism"], np.float64)], axis=1)
        td = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
        self._c = {"A_d": td(self.A_d), "BD_d": td(self.BD_d), "target": td(self.target),
                   "ac_map": td(self.ac_map), "weather": td(weather)}
        self._model = _lib.BuildingModel(
            n, T, self.episode_len, int(not self.is_continuous_action), float(self.q_rate),
            float(self.error_rate), self.reward_pnorm, self._c["A_d"].data_ptr(), self._c["BD_d"].data_ptr(),
            self._c["target"].data_ptr(), self._c["ac_map"].data_ptr(), self._c["weather"].data_ptr())

        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.state = z(B, n + 4, dtype=torch.float32)
        self.epoch = z(B, dtype=torch.int32)
        self.runs = z(B, dtype=torch.int32)
        self.reward = z(B, dtype=torch.float64)
        self.done = z(B, dtype=torch.uint8)
        self.stats = z(_lib.SG_STATS_LEN, dtype=torch.float64) if collect_stats else None
        if self.stats is not None:
            _lib.check(self.lib.sg_stats_reset(self.stats.data_ptr(), self._stream()), "sg_stats_reset")
        self._adtype = torch.float32 if self.is_continuous_action else torch.int64
        self._d_action = z(B, n, dtype=self._adtype)
        self._h_action = torch.zeros(B, n, dtype=self._adtype).pin_memory()
        self._h_obs = torch.zeros(B, n + 4, dtype=torch.float32).pin_memory()
        self._h_reward = torch.zeros(B, dtype=torch.float64).pin_memory()
        self._h_done = torch.zeros(B, dtype=torch.uint8).pin_memory()
        self._flags = _lib.SG_BUILDING_FLAG_AUTORESET if auto_reset else 0
        if not tensor_cores:      # n >= 16 runs on the DMMA cluster kernel unless told otherwise (csrc/building.cu)
            self._flags |= _lib.SG_BUILDING_FLAG_NO_MMA
        self.kernel_launches = 0

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @nvtx_method("sg:building.reset")
    def reset(self, *, seed=None, options: dict | None = None, epochs: np.ndarray | None = None,
              mask: np.ndarray | torch.Tensor | None = None):
        """env.py:314-386.  ``seed``: int (env i gets ``seed + i``), array of per-env seeds, or None
        (random cursors, env.py:340).  ``epochs`` overrides the cursors directly (H6: the RNG-driven
        choice is an input of the engine)."""
        B, n = self.num_envs, self.n
        if epochs is None:
            if seed is None:
                epochs = np.random.default_rng().integers(0, self.length_of_weather - 1, B)
            else:
                seeds = int(seed) + np.arange(B) if np.isscalar(seed) else np.asarray(seed)
                epochs = np.array([seed_to_epoch(int(s), self.episode_len, self.timestep, self.length_of_weather)
                                   for s in seeds])
        d_ep = torch.from_numpy(np.asarray(epochs, np.int32)).to(self.device)
        T0 = None
        if options is not None and "T_initial" in options:
            T0 = torch.from_numpy(np.ascontiguousarray(
                np.broadcast_to(np.asarray(options["T_initial"], np.float64), (B, n)))).to(self.device)
        keep = None
        if mask is not None:        # partial reset: envs outside the mask keep their state
            m = torch.as_tensor(mask, device=self.device).bool()
            keep = (m, self.state.clone(), self.epoch.clone(), self.runs.clone())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_building_reset(C.byref(self._model), B, d_ep.data_ptr(), _lib.ptr(T0),
                                                  self.state.data_ptr(), self.epoch.data_ptr(),
                                                  self.runs.data_ptr(), self._stream()), "sg_building_reset")
        if keep is not None:
            m, st, ep, ru = keep
            self.state.copy_(torch.where(m[:, None], self.state, st))
            self.epoch.copy_(torch.where(m, self.epoch, ep))
            self.runs.copy_(torch.where(m, self.runs, ru))
        self.kernel_launches += 1
        return self.state, {}

    @nvtx_method("sg:building.step")
    def step(self, actions):
        """env.py:205-312 for every env. CUDA tensor in -> CUDA tensors out (async); NumPy in -> NumPy out."""
        B, n = self.num_envs, self.n
        if isinstance(actions, np.ndarray):
            self._h_action.numpy()[...] = actions.reshape(B, n)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.sg_building_step_host(
                    C.byref(self._model), B, self._h_action.data_ptr(), self._h_obs.data_ptr(),
                    self._h_reward.data_ptr(), self._h_done.data_ptr(), self._d_action.data_ptr(),
                    self.state.data_ptr(), self.epoch.data_ptr(), self.runs.data_ptr(), self.reward.data_ptr(),
                    self.done.data_ptr(), self._flags, _lib.ptr(self.stats), self._stream()), "sg_building_step_host")
            self.kernel_launches += 1
            d = self._h_done.numpy().astype(bool)
            return self._h_obs.numpy().copy(), self._h_reward.numpy().copy(), d, d.copy(), {}
        if actions.dtype != self._adtype or tuple(actions.shape) != (B, n) or actions.device != self.device:
            raise ValueError(f"actions must be a {self._adtype} tensor of shape {(B, n)} on {self.device}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_building_step(
                C.byref(self._model), B, actions.contiguous().data_ptr(), self.state.data_ptr(),
                self.epoch.data_ptr(), self.runs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
                self._flags, _lib.ptr(self.stats), self._stream()), "sg_building_step")
        self.kernel_launches += 1
        d = self.done.view(torch.bool)                    # zero-copy (no conversion kernel on the step path)
        return self.state, self.reward, d, d, {}          # terminated == truncated == done (env.py:312)

    @nvtx_method("sg:building.rollout")
    def rollout(self, actions: torch.Tensor, keep_obs: bool = True):
        """``T`` fused steps in one launch. ``actions``: [T, B, n]. Returns (obs [T,B,n+4] or None,
        reward [T,B], done [T,B])."""
        T, B, n = actions.shape
        assert (B, n) == (self.num_envs, self.n) and actions.dtype == self._adtype
        obs = torch.empty(T, B, n + 4, dtype=torch.float32, device=self.device) if keep_obs else None
        rew = torch.empty(T, B, dtype=torch.float64, device=self.device)
        done = torch.empty(T, B, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_building_rollout(
                C.byref(self._model), B, T, actions.contiguous().data_ptr(), self.state.data_ptr(),
                self.epoch.data_ptr(), self.runs.data_ptr(), _lib.ptr(obs), rew.data_ptr(), done.data_ptr(),
                self._flags, _lib.ptr(self.stats), self._stream()), "sg_building_rollout")
        self.kernel_launches += 1
        return obs, rew, done

    def capture_graph(self, actions):
        """One ``step(actions)`` (CUDA tensor, read in place) as a CUDA graph; see ``envs/_graph.py``."""
        from .._graph import capture_step_graph
        return capture_step_graph(self, actions)

    def close(self) -> None:
        pass


class BuildingEnv:
    """The reference's single-env Gymnasium interface (env.py:16-434) on a 1-env batch."""

    def __init__(self, parameters: dict[str, Any] | str, device=None):
        self._b = BatchedBuildingEnv(parameters, 1, device)
        self.parameters = parameters = self._b.parameters
        for k in ("n", "target", "ac_map", "maxpower", "temp_range", "reward_pnorm", "is_continuous_action",
                  "timestep", "episode_len", "length_of_weather", "q_rate", "error_rate", "A_d", "BD_d",
                  "Qlow", "Qhigh", "low", "high"):
            setattr(self, k, getattr(self._b, k))
        self.action_space = self._b.single_action_space
        self.observation_space = self._b.single_observation_space
        self._reward_breakdown = {"comfort_level": 0.0, "power_consumption": 0.0}
        self.np_random = np.random.default_rng()

    @property
    def epoch(self) -> int:
        return int(self._b.epoch[0])

    # the arrays the reference's BuildingEnv keeps as attributes and its MPC agent reads (env.py:150-160, 186-187;
    # algorithms/building/mpc_controller.py:33-37, 50-52)
    @property
    def B_d(self) -> np.ndarray:
        return self.BD_d

    @property
    def out_temp(self) -> np.ndarray:
        return np.asarray(self.parameters["out_temp"], np.float64)

    @property
    def ground_temp(self) -> np.ndarray:
        return np.asarray(self.parameters["ground_temp"], np.float64)

    @property
    def ghi(self) -> np.ndarray:
        return np.asarray(self.parameters["ghi"], np.float64)

    @property
    def metabolism(self) -> np.ndarray:
        return np.asarray(self.parameters["metabolism"], np.float64)

    def _info(self) -> dict:
        return {"zone_temperature": self.state[: self.n].astype(np.float64),
                "reward_breakdown": self._reward_breakdown}

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
            epochs = None
        else:   # env.py:340
            epochs = np.array([self.np_random.integers(low=0, high=self.length_of_weather - 1)])
        obs, _ = self._b.reset(seed=seed, options=options, epochs=epochs)
        self.state = obs.cpu().numpy()[0].copy()
        for k in self._reward_breakdown:
            self._reward_breakdown[k] = 0.0
        return self.state, self._info()

    def step(self, action):
        dt = np.float32 if self.is_continuous_action else np.int64
        action = np.asarray(action, dt).reshape(self.n)
        obs, rew, done, _, _ = self._b.step(action.reshape(1, self.n))
        self.state = obs[0]
        # reward breakdown (env.py:276-280): the power term is the reference's own expression on the action, the
        # comfort term is the rest of the reward the kernel returned (reward = -(power + comfort))
        a = action if self.is_continuous_action else (action + self.Qlow * DISCRETE_LENGTH) / DISCRETE_LENGTH   # env.py:234-235
        power = float(np.linalg.norm(a, self.reward_pnorm) * self.q_rate)
        self._reward_breakdown["power_consumption"] -= power
        self._reward_breakdown["comfort_level"] -= -float(rew[0]) - power
        return self.state, float(rew[0]), bool(done[0]), bool(done[0]), self._info()

    def close(self) -> None:
        pass
