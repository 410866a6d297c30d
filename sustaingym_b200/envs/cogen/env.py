"""Batched CogenEnv on the B200 engine (single- and multi-agent views).

``BatchedCogenEnv`` advances ``num_envs`` copies of the reference's ``CogenEnv``
(``sustaingym/envs/cogen/env.py``) per call; the plant model (the reference's ONNX MLP ensemble,
repacked in ``sustaingym_b200/data/cogen_plant.npz``) runs in ``sg_cogen_step``.  ``CogenEnv`` /
``MultiAgentCogenEnv`` are the ``num_envs = 1`` Gymnasium / PettingZoo views with the reference's
Dict action/observation structure.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any

import numpy as np
import torch

from ... import _lib, spaces
from .._state import StateDictMixin, nvtx_method
from ...data.cogen_ambients import AMBIENT_COLS, real_ambients, synthetic_ambients  # noqa: F401

PLANT_NPZ = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                         "data", "cogen_plant.npz")
# batched action row = model inputs 3..17 (model.json order)
ACTION_KEYS = ("GT1_PAC_FFU", "GT1_EVC_FFU", "GT1_PWR", "GT2_PAC_FFU", "GT2_EVC_FFU", "GT2_PWR",
               "GT3_PAC_FFU", "GT3_EVC_FFU", "GT3_PWR", "HR1_HPIP_M_PROC", "HR2_HPIP_M_PROC",
               "HR3_HPIP_M_PROC", "ST_PWR", "IPPROC_M", "CT_NrBays")
AGENTS = ("GT1", "GT2", "GT3", "ST")
AGENT_ACTION_KEYS = {          # multiagent_env.py:50-55
    "GT1": ["GT1_PWR", "GT1_PAC_FFU", "GT1_EVC_FFU", "HR1_HPIP_M_PROC"],
    "GT2": ["GT2_PWR", "GT2_PAC_FFU", "GT2_EVC_FFU", "HR2_HPIP_M_PROC"],
    "GT3": ["GT3_PWR", "GT3_PAC_FFU", "GT3_EVC_FFU", "HR3_HPIP_M_PROC"],
    "ST": ["ST_PWR", "IPPROC_M", "CT_NrBays"],
}


class PlantWeights:
    """Device copy of the packed plant model (shared by every env on that device)."""

    def __init__(self, device, path: str = PLANT_NPZ):
        z = np.load(path)
        pad = lambda a, dt: np.ascontiguousarray(np.pad(a, ((0, 0), (0, 16 - a.shape[1])), constant_values=(-1 if dt == np.int32 else 0)).astype(dt))  # noqa: E731
        self.in_ids = [str(s) for s in z["in_ids"]]
        self.in_min, self.in_max = z["in_min"], z["in_max"]
        self.out_min, self.out_max = z["out_min"], z["out_max"]
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        self.blob = t(z["blob"].astype(np.float32))
        self.offsets = t(z["offsets"].astype(np.int64))
        self.meta = t(z["meta"].astype(np.int32))
        self.feats = t(pad(z["feats"], np.int32))
        self.mu = t(pad(z["mu"], np.float32))
        self.sinv = t(pad(z["sinv"], np.float32))
        self.b3 = t(z["b3"].astype(np.float32))
        self.tc_blob = self.tc_offsets = None        # filled on first use (sg_cogen_prepare_tc)

    _cache: dict = {}

    @classmethod
    def shared(cls, device) -> "PlantWeights":
        """One copy per device: every env instance (and every CTA) reads the same L2-resident weights."""
        key = str(device)
        if key not in cls._cache:
            cls._cache[key] = cls(device)
        return cls._cache[key]


def make_spaces(weights: PlantWeights, forecast_horizon: int):
    """Action / observation spaces exactly as cogen/env.py:114-143 builds them."""
    lo = dict(zip(weights.in_ids, weights.in_min))
    hi = dict(zip(weights.in_ids, weights.in_max))
    box = lambda k: spaces.Box(low=lo[k], high=hi[k], shape=(1,), dtype=np.float32)  # noqa: E731
    act = {}
    for i in (1, 2, 3):
        act[f"GT{i}_PWR"] = box(f"GT{i}_PWR")
        act[f"GT{i}_PAC_FFU"] = spaces.Discrete(2)
        act[f"GT{i}_EVC_FFU"] = spaces.Discrete(2)
        act[f"HR{i}_HPIP_M_PROC"] = box(f"HR{i}_HPIP_M_PROC")
    act["ST_PWR"] = box("ST_PWR")
    act["IPPROC_M"] = box("IPPROC_M")
    act["CT_NrBays"] = spaces.Discrete(12, start=1)
    action_space = spaces.Dict(act)
    k1 = forecast_horizon + 1
    fbox = lambda l, h: spaces.Box(low=l, high=h, shape=(k1,), dtype=np.float32)  # noqa: E731
    observation_space = spaces.Dict({
        "Time": spaces.Box(low=0, high=1, shape=(1,), dtype=np.float32),
        "Prev_Action": action_space,
        "TAMB": fbox(lo["TAMB"], hi["TAMB"]), "PAMB": fbox(lo["PAMB"], hi["PAMB"]),
        "RHAMB": fbox(lo["RHAMB"], hi["RHAMB"]), "Target_Power": fbox(0, 700), "Target_Steam": fbox(0, 1300),
        "Energy_Price": fbox(0, 1500), "Gas_Price": fbox(0, 7),
    })
    return action_space, observation_space


def action_dict_to_row(action: dict) -> np.ndarray:
    return np.array([np.asarray(action[k], np.float32).reshape(-1)[0] for k in ACTION_KEYS], np.float32)


def action_row_to_dict(row: np.ndarray) -> dict:
    out = {}
    for i, k in enumerate(ACTION_KEYS):
        if k.endswith("_FFU") or k == "CT_NrBays":
            out[k] = np.int64(row[i])
        else:
            out[k] = np.array([row[i]], np.float32)
    return out


class BatchedCogenEnv(StateDictMixin):
    """Constructor mirrors cogen/env.py:75-82.  ``ambients`` ([days, 96, 7]) overrides the table; by default it
    is the reference's own xlsx-derived table for ``renewables_magnitude`` (packaged, ``data/cogen_ambients.py``)."""

    _STATE_TENSORS = ('prev_action', 'day_t', 'obs', 'reward', 'agent_reward', 'done', 'Y', 'stats')          # checkpointed by state_dict() (envs/_state.py)

    def __init__(self, num_envs: int = 1, ambients: np.ndarray | None = None, renewables_magnitude: float = 0.,
                 ramp_penalty: float = 2., supply_imbalance_penalty: float = 1000,
                 constraint_violation_penalty: float = 1000, forecast_horizon: int = 3,
                 forecast_noise_std: float = 0.0, device=None, collect_stats: bool = False, keep_outputs: bool = False,
                 tensor_cores: bool = True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedCogenEnv needs a CUDA device (no CPU fallback)")
        self.device = dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_envs = B = int(num_envs)
        self.ramp_penalty, self.supply_imbalance_penalty = ramp_penalty, supply_imbalance_penalty
        self.constraint_violation_penalty = constraint_violation_penalty
        self.forecast_horizon, self.forecast_noise_std = forecast_horizon, forecast_noise_std
        if ambients is None:      # the reference's table for this renewables_magnitude (cogen/env.py:89)
            ambients = real_ambients(renewables_magnitude)
        self.ambients = np.ascontiguousarray(ambients, np.float64)
        self.n_days, self.timesteps_per_day = self.ambients.shape[0], self.ambients.shape[1]
        assert self.timesteps_per_day == 96 and self.ambients.shape[2] == 7
        assert 0 <= forecast_horizon < self.timesteps_per_day - 1       # cogen/env.py:92
        self.weights = PlantWeights.shared(dev)
        self.single_action_space, self.single_observation_space = make_spaces(self.weights, forecast_horizon)
        self.obs_dim = 7 * (forecast_horizon + 1) + 33
        self._amb = torch.from_numpy(self.ambients).to(dev)
        w = self.weights
        self._model = _lib.CogenModel(
            23, forecast_horizon, self.n_days, 0, float(ramp_penalty), float(supply_imbalance_penalty),
            float(constraint_violation_penalty), w.blob.data_ptr(), w.offsets.data_ptr(), w.meta.data_ptr(),
            w.feats.data_ptr(), w.mu.data_ptr(), w.sinv.data_ptr(), w.b3.data_ptr(), self._amb.data_ptr(), None, None)
        if tensor_cores:
            # layer-2 weights pre-split into tf32 (hi, lo) pairs in the tcgen05 operand layout (3xTF32 path)
            if w.tc_blob is None:
                meta = w.meta.cpu().numpy()
                sizes = (meta[:, 1] // 16) * (272 + 32 * meta[:, 2])
                offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
                w.tc_offsets = torch.from_numpy(offs[:-1].copy()).to(dev)
                w.tc_blob = torch.zeros(int(offs[-1]), dtype=torch.float32, device=dev)
                self._model.tc_offsets = w.tc_offsets.data_ptr()
                with torch.cuda.device(dev):
                    _lib.check(self.lib.sg_cogen_prepare_tc(C.byref(self._model), w.tc_blob.data_ptr(),
                                                            torch.cuda.current_stream(dev).cuda_stream), "sg_cogen_prepare_tc")
            self._model.tc_offsets = w.tc_offsets.data_ptr()
            self._model.tc_blob = w.tc_blob.data_ptr()
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.prev_action = z(B, 15, dtype=torch.float32)
        self.day_t = z(B, 2, dtype=torch.int32)
        self.obs = z(B, self.obs_dim, dtype=torch.float32)
        self.reward = z(B, dtype=torch.float64)
        self.agent_reward = z(B, 4, dtype=torch.float64)
        self.done = z(B, dtype=torch.uint8)
        self._trunc = z(B, dtype=torch.bool)              # truncated is always False (cogen/env.py:384)
        self.scratch = z(_lib.cogen_scratch_floats(B), dtype=torch.float32)
        self.Y = z(B, 29, dtype=torch.float32) if keep_outputs else None
        self.stats = z(_lib.SG_STATS_LEN, dtype=torch.float64) if collect_stats else None
        if self.stats is not None:
            _lib.check(self.lib.sg_stats_reset(self.stats.data_ptr(), self._stream()), "sg_stats_reset")
        self._gen = torch.Generator(device=dev)
        self.kernel_launches = 0

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def sample_actions(self, generator: np.random.Generator) -> np.ndarray:
        """[B,15] rows drawn like ``action_space.sample()`` (cogen/env.py:222-223)."""
        w, B = self.weights, self.num_envs
        a = generator.uniform(w.in_min[3:], w.in_max[3:], (B, 15)).astype(np.float32)
        for c in (0, 1, 3, 4, 6, 7):
            a[:, c] = generator.integers(0, 2, B)
        a[:, 14] = generator.integers(1, 13, B)
        return a

    def plant(self, X: torch.Tensor) -> torch.Tensor:
        """The plant model alone: X [N,18] f32 -> Y [N,29] f32 (what ORT computes at cogen/env.py:305)."""
        N = X.shape[0]
        Y = torch.empty(N, 29, dtype=torch.float32, device=self.device)
        scratch = torch.empty(_lib.cogen_scratch_floats(N), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_cogen_plant(C.byref(self._model), N, X.contiguous().data_ptr(), Y.data_ptr(),
                                               scratch.data_ptr(), self._stream()), "sg_cogen_plant")
        self.kernel_launches += 2
        return Y

    @nvtx_method("sg:cogen.reset")
    def reset(self, *, seed=None, options=None, days: np.ndarray | None = None, init_actions: np.ndarray | None = None,
              mask: np.ndarray | torch.Tensor | None = None):
        """cogen/env.py:186-230. ``seed``: int (env i gets seed+i) or per-env array -> day = seed % n_days
        and the initial Prev_Action is drawn from a generator seeded with it; or pass ``days`` /
        ``init_actions`` explicitly (H6: RNG-driven choices are inputs of the engine)."""
        B = self.num_envs
        if seed is None:
            seeds = np.random.default_rng().integers(0, 2 ** 31, B)
            if days is None:
                days = np.random.default_rng().integers(0, self.n_days - 1, B)   # cogen/env.py:214
        else:
            seeds = int(seed) + np.arange(B) if np.isscalar(seed) else np.asarray(seed)
        if days is None:
            days = seeds % self.n_days                                         # cogen/env.py:216
        if init_actions is None:
            init_actions = self.sample_actions(np.random.default_rng(int(seeds[0])))
        d_days = torch.from_numpy(np.asarray(days, np.int32)).to(self.device)
        d_act = torch.from_numpy(np.ascontiguousarray(init_actions, np.float32)).to(self.device)
        keep = None
        if mask is not None:        # partial reset: envs outside the mask keep their state
            m = torch.as_tensor(mask, device=self.device).bool()
            keep = (m, self.prev_action.clone(), self.day_t.clone(), self.obs.clone())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_cogen_reset(C.byref(self._model), B, d_days.data_ptr(), d_act.data_ptr(),
                                               self.prev_action.data_ptr(), self.day_t.data_ptr(), self.obs.data_ptr(),
                                               self._stream()), "sg_cogen_reset")
        if keep is not None:
            m, pa, dt, ob = keep
            self.prev_action.copy_(torch.where(m[:, None], self.prev_action, pa))
            self.day_t.copy_(torch.where(m[:, None], self.day_t, dt))
            self.obs.copy_(torch.where(m[:, None], self.obs, ob))
        self.kernel_launches += 1
        return self.obs, {}

    @nvtx_method("sg:cogen.step")
    def step(self, actions, noise: torch.Tensor | None = None):
        """cogen/env.py:355-387. ``actions``: [B,15] f32 CUDA tensor (or NumPy -> NumPy out).  ``noise``:
        optional [B,k,7] f64 standard-normal draws (scaled here by forecast_noise_std); when the env
        was built with forecast_noise_std > 0 and none is given, torch draws it on the device."""
        B, k = self.num_envs, self.forecast_horizon
        host = isinstance(actions, np.ndarray)
        a = torch.from_numpy(np.ascontiguousarray(actions, np.float32)).to(self.device) if host else actions.contiguous()
        if noise is None and self.forecast_noise_std > 0 and k > 0:
            noise = torch.randn(B, k, 7, dtype=torch.float64, device=self.device, generator=self._gen)
        d_noise = None
        if noise is not None and k > 0:
            d_noise = (noise.to(self.device, torch.float64) * self.forecast_noise_std).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_cogen_step(
                C.byref(self._model), B, a.data_ptr(), self.prev_action.data_ptr(), self.day_t.data_ptr(),
                _lib.ptr(d_noise), self.obs.data_ptr(), self.reward.data_ptr(), self.agent_reward.data_ptr(),
                self.done.data_ptr(), _lib.ptr(self.Y), self.scratch.data_ptr(), 0, _lib.ptr(self.stats),
                self._stream()), "sg_cogen_step")
        self.kernel_launches += 2
        info = {"agent_rewards": self.agent_reward}
        if host:
            d = self.done.cpu().numpy().astype(bool)
            return self.obs.cpu().numpy(), self.reward.cpu().numpy(), d, np.zeros(B, bool), \
                {"agent_rewards": self.agent_reward.cpu().numpy()}
        return self.obs, self.reward, self.done.view(torch.bool), self._trunc, info

    def close(self) -> None:
        pass


def _unflatten_obs(flat: np.ndarray, k: int) -> dict:
    """Inverse of the sorted-key flatten for one env: returns the reference's Dict observation."""
    k1 = k + 1
    pos, out = 0, {}
    for key in ("Energy_Price", "Gas_Price", "PAMB"):
        out[key] = flat[pos:pos + k1].astype(np.float32); pos += k1
    pa = flat[pos:pos + 32]; pos += 32
    prev, q = {}, 0
    prev["CT_NrBays"] = np.int64(int(np.argmax(pa[q:q + 12])) + 1); q += 12
    for i in (1, 2, 3):
        prev[f"GT{i}_EVC_FFU"] = np.int64(int(np.argmax(pa[q:q + 2]))); q += 2
        prev[f"GT{i}_PAC_FFU"] = np.int64(int(np.argmax(pa[q:q + 2]))); q += 2
        prev[f"GT{i}_PWR"] = pa[q:q + 1].astype(np.float32); q += 1
    for i in (1, 2, 3):
        prev[f"HR{i}_HPIP_M_PROC"] = pa[q:q + 1].astype(np.float32); q += 1
    prev["IPPROC_M"] = pa[q:q + 1].astype(np.float32); q += 1
    prev["ST_PWR"] = pa[q:q + 1].astype(np.float32); q += 1
    out["Prev_Action"] = prev
    for key in ("RHAMB", "TAMB", "Target_Power", "Target_Steam"):
        out[key] = flat[pos:pos + k1].astype(np.float32); pos += k1
    out["Time"] = flat[pos:pos + 1].astype(np.float32)
    return out


class CogenEnv:
    """The reference's single-env Gymnasium interface (cogen/env.py:18-387): Dict actions/observations."""

    def __init__(self, renewables_magnitude: float = 0., ramp_penalty: float = 2., supply_imbalance_penalty: float = 1000,
                 constraint_violation_penalty: float = 1000, forecast_horizon: int = 3, forecast_noise_std: float = 0.0,
                 ambients: np.ndarray | None = None, device=None):
        self._b = BatchedCogenEnv(1, ambients, renewables_magnitude, ramp_penalty, supply_imbalance_penalty,
                                  constraint_violation_penalty, forecast_horizon, forecast_noise_std, device,
                                  keep_outputs=True)
        self.ramp_penalty, self.supply_imbalance_penalty = ramp_penalty, supply_imbalance_penalty
        self.action_space = self._b.single_action_space
        self.observation_space = self._b.single_observation_space
        self.forecast_horizon, self.forecast_noise_std = forecast_horizon, forecast_noise_std
        self.n_days, self.timesteps_per_day = self._b.n_days, 96
        self.np_random = np.random.default_rng()
        self.t = 0

    def reset(self, *, seed: int | None = None, options: dict | None = None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        day = int(self.np_random.integers(low=0, high=self.n_days - 1)) if seed is None else seed % self.n_days
        self.action_space.seed(seed)
        init = action_dict_to_row(self.action_space.sample())[None]
        obs, _ = self._b.reset(seed=0 if seed is None else seed, days=np.array([day]), init_actions=init)
        self.current_day, self.t = day, 0
        self.obs = _unflatten_obs(obs.cpu().numpy()[0], self.forecast_horizon)
        return self.obs, {"Operating constraint violation": None, "Demand constraint violation": None}

    def step(self, action: dict[str, Any]):
        noise = None
        if self.forecast_noise_std > 0 and self.forecast_horizon > 0:        # cogen/env.py:161 (env's own RNG)
            noise = torch.from_numpy(self.np_random.normal(size=(1, self.forecast_horizon, 7)))
        prev_obs = self.obs
        obs, rew, done, _, info = self._b.step(action_dict_to_row(action)[None], noise)
        self.t += 1
        self.obs = _unflatten_obs(obs[0], self.forecast_horizon)
        ar = info["agent_rewards"][0]
        self.current_info = {"agent_rewards": {a: float(ar[i]) for i, a in enumerate(AGENTS)}}
        self.current_info.update(self._reward_breakdown(prev_obs, action, ar))
        return self.obs, float(rew[0]), bool(done[0]), False, self.current_info

    def _reward_breakdown(self, prev_obs: dict, action: dict, agent_rewards: np.ndarray) -> dict:
        """The reference's info dict (cogen/env.py:309-352) from what the kernels emit: fuel costs are plant outputs
        6..8, ramp and non-delivery costs are the reference's own f32 expressions on the action / previous observation
        / plant outputs 27, 28, and each unit's constraint-violation cost is what is left of its reward share
        (multiagent_env.py:97-101: reward_i = -(fuel_i + ramp_i + dyn_cv_i + non_delivery / 4))."""
        Y = self._b.Y[0].cpu().numpy()
        rp = np.float32(self.ramp_penalty)
        fuel = {"GT1": Y[6], "GT2": Y[7], "GT3": Y[8], "ST": 0}
        prev = prev_obs["Prev_Action"]
        pw = {"GT1": "GT1_PWR", "GT2": "GT2_PWR", "GT3": "GT3_PWR", "ST": "ST_PWR"}
        ramp = {a: rp * np.abs(np.float32(np.asarray(action[k]).reshape(-1)[0]) - np.float32(np.asarray(prev[k]).reshape(-1)[0]))
                for a, k in pw.items()}
        steam_penalty = np.maximum(np.float32(0), prev_obs["Target_Steam"][0] - Y[28])
        energy_penalty = np.maximum(np.float32(0), prev_obs["Target_Power"][0] - Y[27])
        non_delivery = np.float32(self.supply_imbalance_penalty) * (steam_penalty + energy_penalty)
        dyn = {a: max(0.0, -float(agent_rewards[i]) - float(fuel[a]) - float(ramp[a]) - float(non_delivery) / 4)
               for i, a in enumerate(AGENTS)}
        return {"fuel_costs": fuel, "ramp_costs": ramp, "dyn_cv_costs": dyn, "non_delivery_cost": non_delivery}

    def close(self) -> None:
        pass


class MultiAgentCogenEnv:
    """PettingZoo ParallelEnv view (cogen/multiagent_env.py:15-137): agents GT1, GT2, GT3, ST; every
    agent receives the same flattened observation; rewards split per unit."""

    metadata: dict[str, Any] = {}

    def __init__(self, renewables_magnitude: float = 0., ramp_penalty: float = 2., supply_imbalance_penalty: float = 1000,
                 constraint_violation_penalty: float = 1000, forecast_horizon: int = 12, forecast_noise_std: float = 0.1,
                 ambients: np.ndarray | None = None, device=None):
        self.single_env = CogenEnv(renewables_magnitude, ramp_penalty, supply_imbalance_penalty,
                                   constraint_violation_penalty, forecast_horizon, forecast_noise_std, ambients, device)
        self.agents = list(AGENTS)
        self.possible_agents = self.agents
        flat = spaces.flatten_space(self.single_env.observation_space)
        self.observation_spaces = {a: flat for a in self.agents}
        self.agents_to_action_keys = AGENT_ACTION_KEYS
        self.action_spaces = {a: spaces.Dict({k: self.single_env.action_space[k] for k in keys})
                              for a, keys in AGENT_ACTION_KEYS.items()}

    @property
    def num_agents(self) -> int:
        return len(self.agents)

    def _flat(self) -> np.ndarray:
        # the reference's flatten is float64 (int64 one-hots meet f32 boxes); same values
        return self.single_env._b.obs.cpu().numpy()[0].astype(np.float64)

    def step(self, actions: dict[str, dict]):
        action: dict = {}
        for agent in self.agents:
            action |= actions[agent]
        _, _, terminated, truncated, info = self.single_env.step(action)
        flat = self._flat()
        obss, rewards, terms, truncs, infos = {}, {}, {}, {}, {}
        for agent in self.agents:
            obss[agent] = flat
            rewards[agent] = info["agent_rewards"][agent]
            terms[agent], truncs[agent], infos[agent] = terminated, truncated, {}
        if terminated or truncated:
            self.agents = []
        return obss, rewards, terms, truncs, infos

    def reset(self, seed: int | None = None, options: dict | None = None):
        _, info = self.single_env.reset(seed=seed, options=options)
        flat = self._flat()
        self.agents = self.possible_agents[:]
        return {a: flat for a in self.agents}, {a: info for a in self.agents}

    def observation_space(self, agent: str):
        return self.observation_spaces[agent]

    def action_space(self, agent: str):
        return self.action_spaces[agent]

    def render(self) -> None:
        pass

    def close(self) -> None:
        self.single_env.close()
