"""Batched ElectricityMarketEnv — SPEC-DEFINED (see SPEC.md next to this file: the reference snapshot
ships no source for this environment, only ``docs/electricitymarketenv.md``)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ... import _lib, spaces
from .._state import StateDictMixin, nvtx_method

T_STEPS = 288


def default_fleet(G: int = 33, seed: int = 0):
    """33 units with an RTS-like spread of sizes and costs (builder's parameters, not reference data)."""
    rng = np.random.default_rng(seed)
    cost = np.sort(np.concatenate([rng.uniform(5, 25, G // 3), rng.uniform(25, 60, G // 3),
                                   rng.uniform(60, 250, G - 2 * (G // 3))]))
    cap = np.concatenate([rng.uniform(150, 400, G // 3), rng.uniform(50, 200, G // 3),
                          rng.uniform(10, 100, G - 2 * (G // 3))])
    return cost, cap


def synthetic_load_moer(n_days: int, k: int = 1, seed: int = 0):
    """Day tables [D, 288+k]: load (MW) with a double-peak daily shape sized to the default fleet, MOER
    (kg CO2/MWh) in the range of the packaged SGIP tables (0-660)."""
    rng = np.random.default_rng(seed)
    t = np.arange(T_STEPS + k) / T_STEPS
    shape = 0.62 + 0.18 * np.sin(2 * np.pi * (t - 0.35)) + 0.10 * np.sin(4 * np.pi * (t - 0.2))
    load = 4300 * shape[None, :] * rng.uniform(0.85, 1.1, (n_days, 1)) + rng.normal(0, 40, (n_days, T_STEPS + k))
    moer = np.clip(330 + 220 * np.sin(2 * np.pi * (t[None, :] - 0.6 + rng.uniform(-0.1, 0.1, (n_days, 1))))
                   + rng.normal(0, 25, (n_days, T_STEPS + k)), 0, 660)
    return load, moer


class BatchedElectricityMarketEnv(StateDictMixin):

    _STATE_TENSORS = ('state', 'day_t', 'prev_action', 'obs', 'reward', 'done', 'info', 'stats')          # checkpointed by state_dict() (envs/_state.py)
    def __init__(self, num_envs: int = 1, load: np.ndarray | None = None, moer: np.ndarray | None = None,
                 gen_cost: np.ndarray | None = None, gen_cap: np.ndarray | None = None, k: int = 1,
                 use_intermediate_rewards: bool = True, e_max: float = 80.0, e_init: float = 40.0,
                 p_max: float = 20.0, eta_c: float = 0.95, eta_d: float = 0.95, price_cap: float = 1000.0,
                 p_co2: float = 0.03085, terminal_price: float = 50.0, device=None, collect_stats: bool = False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedElectricityMarketEnv needs a CUDA device (no CPU fallback)")
        self.device = dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_envs = B = int(num_envs)
        self.k = k
        if gen_cost is None:
            gen_cost, gen_cap = default_fleet()
        if load is None:
            load, moer = synthetic_load_moer(32, k)
        order = np.argsort(gen_cost, kind="stable")
        self.gen_cost, self.gen_cap = np.asarray(gen_cost, float)[order], np.asarray(gen_cap, float)[order]
        self.load, self.moer = np.ascontiguousarray(load, np.float64), np.ascontiguousarray(moer, np.float64)
        assert self.load.shape == self.moer.shape and self.load.shape[1] == T_STEPS + k
        self.n_days = self.load.shape[0]
        self.obs_dim = 6 + 4 * k
        td = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(dev)  # noqa: E731
        self._c = {"cost": td(self.gen_cost), "cap": td(self.gen_cap), "load": td(self.load), "moer": td(self.moer)}
        self._model = _lib.MarketModel(len(self.gen_cost), k, self.n_days, int(use_intermediate_rewards), e_max, e_init,
                                       p_max, eta_c, eta_d, price_cap, p_co2, terminal_price,
                                       self._c["cost"].data_ptr(), self._c["cap"].data_ptr(),
                                       self._c["load"].data_ptr(), self._c["moer"].data_ptr())
        self.single_action_space = spaces.Box(0.0, price_cap, shape=(2 * k,), dtype=np.float32)
        self.single_observation_space = spaces.Box(-np.inf, np.inf, shape=(self.obs_dim,), dtype=np.float32)
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.state = z(B, 4, dtype=torch.float64)
        self.day_t = z(B, 2, dtype=torch.int32)
        self.prev_action = z(B, 2 * k, dtype=torch.float32)
        self.obs = z(B, self.obs_dim, dtype=torch.float32)
        self.reward = z(B, dtype=torch.float64)
        self.done = z(B, dtype=torch.uint8)
        self._trunc = z(B, dtype=torch.bool)
        self.info = z(B, 3, dtype=torch.float64)
        self.stats = z(_lib.SG_STATS_LEN, dtype=torch.float64) if collect_stats else None
        if self.stats is not None:
            _lib.check(self.lib.sg_stats_reset(self.stats.data_ptr(), self._stream()), "sg_stats_reset")
        self.kernel_launches = 0

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @nvtx_method("sg:market.reset")
    def reset(self, *, seed=None, options=None, days: np.ndarray | None = None, mask=None):
        B = self.num_envs
        if days is None:
            if seed is None:
                days = np.random.default_rng().integers(0, self.n_days, B)
            else:
                seeds = int(seed) + np.arange(B) if np.isscalar(seed) else np.asarray(seed)
                days = seeds % self.n_days
        d = torch.from_numpy(np.asarray(days, np.int32)).to(self.device)
        keep = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).bool()
            keep = (m, self.state.clone(), self.day_t.clone(), self.prev_action.clone(), self.obs.clone())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_market_reset(C.byref(self._model), B, d.data_ptr(), self.state.data_ptr(),
                                                self.day_t.data_ptr(), self.prev_action.data_ptr(), self.obs.data_ptr(),
                                                self._stream()), "sg_market_reset")
        if keep is not None:
            m, st, dt, pa, ob = keep
            self.state.copy_(torch.where(m[:, None], self.state, st)); self.day_t.copy_(torch.where(m[:, None], self.day_t, dt))
            self.prev_action.copy_(torch.where(m[:, None], self.prev_action, pa)); self.obs.copy_(torch.where(m[:, None], self.obs, ob))
        self.kernel_launches += 1
        return self.obs, {}

    @nvtx_method("sg:market.step")
    def step(self, actions):
        B = self.num_envs
        host = isinstance(actions, np.ndarray)
        a = torch.from_numpy(np.ascontiguousarray(actions, np.float32)).to(self.device) if host else actions.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sg_market_step(C.byref(self._model), B, a.data_ptr(), self.state.data_ptr(),
                                               self.day_t.data_ptr(), self.prev_action.data_ptr(), self.obs.data_ptr(),
                                               self.reward.data_ptr(), self.done.data_ptr(), self.info.data_ptr(),
                                               _lib.ptr(self.stats), self._stream()), "sg_market_step")
        self.kernel_launches += 1
        info = {"price": self.info[:, 0], "dispatch": self.info[:, 1], "energy": self.info[:, 2]}
        if host:
            d = self.done.cpu().numpy().astype(bool)
            return self.obs.cpu().numpy(), self.reward.cpu().numpy(), d, np.zeros(B, bool), \
                {k_: v.cpu().numpy() for k_, v in info.items()}
        return self.obs, self.reward, self.done.view(torch.bool), self._trunc, info

    def capture_graph(self, actions):
        """One ``step(actions)`` (CUDA tensor, read in place) as a CUDA graph; see ``envs/_graph.py``."""
        from .._graph import capture_step_graph
        return capture_step_graph(self, actions)

    def close(self) -> None:
        pass


class ElectricityMarketEnv:
    """Single-env Gymnasium view (docs/electricitymarketenv.md 'Custom RL Loop')."""

    def __init__(self, month: str | None = None, use_intermediate_rewards: bool = True, **kw):
        self.month = month
        self._b = BatchedElectricityMarketEnv(1, use_intermediate_rewards=use_intermediate_rewards, **kw)
        self.action_space, self.observation_space = self._b.single_action_space, self._b.single_observation_space

    def reset(self, *, seed: int | None = None, options=None):
        obs, _ = self._b.reset(seed=seed)
        self.action_space.seed(seed)
        return obs.cpu().numpy()[0], {}

    def step(self, action):
        obs, r, term, trunc, info = self._b.step(np.asarray(action, np.float32).reshape(1, -1))
        return obs[0], float(r[0]), bool(term[0]), False, {k_: float(v[0]) for k_, v in info.items()}

    def close(self) -> None:
        pass


class DiscreteMarketWrapper:
    """3 actions: 0 charge (bid the price cap to buy), 1 do nothing, 2 discharge (ask 0 to sell)."""

    def __init__(self, env: ElectricityMarketEnv):
        self.env = env
        self.action_space = spaces.Discrete(3)
        self.observation_space = env.observation_space
        self._cap = float(env._b.single_action_space.high.max())

    def action(self, a: int) -> np.ndarray:
        k = self.env._b.k
        out = np.zeros(2 * k, np.float32)
        out[k:] = self._cap                      # never discharge unless asked
        if a == 0:
            out[:k] = self._cap
        elif a == 2:
            out[k:] = 0.0
        return out

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, a):
        return self.env.step(self.action(int(a)))
