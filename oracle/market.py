"""TEST INFRASTRUCTURE — CPU oracle for ElectricityMarketEnv.  SPEC-DEFINED, NOT REFERENCE-DEFINED.

The reference snapshot (chrisyeh96/sustaingym @ eac2a4d) ships NO source for this environment
(``sustaingym/__init__.py:9-12`` has the registration commented out; only
``docs/electricitymarketenv.md`` remains — SURVEY.md F2).  There is therefore nothing to be in
parity with; this module is the executable statement of the spec the batched kernel implements
(``sustaingym_b200/envs/electricitymarket/SPEC.md``), which keeps what the docs page states
(docs/electricitymarketenv.md:9-27): T = 288 five-minute steps, tau = 5/60 h, 33 generators bidding
true cost, one 80 MWh battery bidding (a^c, a^d), observation ``(t, e, a(t-1), x_{t-1}, p_{t-1},
l_{t-1}, lhat, m_{t-1}, mhat)``, reward ``p_t x_t + P_CO2 m_t x_t - c_T`` with a terminal cost only at
t = T, optional terminal-only rewards — and replaces the multi-period network-constrained SCED (no
data for it in the snapshot) by a single-bus, single-period merit-order clearing.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

T_STEPS = 288
TAU = 5.0 / 60.0


@dataclass
class MarketParams:
    gen_cost: np.ndarray        # [G] $/MWh, ascending (merit order)
    gen_cap: np.ndarray         # [G] MW
    e_max: float = 80.0         # MWh
    e_init: float = 40.0        # MWh
    p_max: float = 20.0         # MW charge/discharge limit
    eta_c: float = 0.95
    eta_d: float = 0.95
    price_cap: float = 1000.0   # $/MWh when generation is short
    p_co2: float = 0.03085      # $/kg CO2 (30.85 $/t, the EV env's carbon price)
    terminal_price: float = 50.0
    k: int = 1
    use_intermediate_rewards: bool = True


def default_fleet(G: int = 33, seed: int = 0):
    """33 units with RTS-like spread of sizes and costs (builder's parameters, not reference data)."""
    rng = np.random.default_rng(seed)
    cost = np.sort(np.concatenate([rng.uniform(5, 25, G // 3), rng.uniform(25, 60, G // 3),
                                   rng.uniform(60, 250, G - 2 * (G // 3))]))
    cap = np.concatenate([rng.uniform(150, 400, G // 3), rng.uniform(50, 200, G // 3),
                          rng.uniform(10, 100, G - 2 * (G // 3))])
    return cost, cap


def clear(p: MarketParams, load, e, a_c, a_d):
    """Single-bus merit-order clearing. All inputs [B] f64. Returns (x_d, x_c, price) in MW, MW, $/MWh."""
    cost, cap = p.gen_cost, p.gen_cap
    cum = np.cumsum(cap)
    q_d = np.minimum(p.p_max, e * p.eta_d / TAU)                     # MW the battery can deliver this step
    q_c = np.minimum(p.p_max, (p.e_max - e) / (p.eta_c * TAU))       # MW it can absorb
    q_d, q_c = np.maximum(q_d, 0.0), np.maximum(q_c, 0.0)
    # generation cheaper than the discharge offer goes first (ties: generators first)
    cheap_d = np.where(cost[None, :] <= a_d[:, None], cap[None, :], 0.0).sum(1)
    x_d = np.clip(load - cheap_d, 0.0, q_d)
    # the battery charges from generation priced at or below its bid, after the load is served
    cheap_c = np.where(cost[None, :] <= a_c[:, None], cap[None, :], 0.0).sum(1)
    x_c = np.where(x_d > 0, 0.0, np.clip(cheap_c - load, 0.0, q_c))
    net = load - x_d + x_c
    idx = np.sum(cum[None, :] < net[:, None] - 1e-12, axis=1)        # marginal generator
    gen_price = np.where(idx < len(cost), cost[np.minimum(idx, len(cost) - 1)], p.price_cap)
    gen_price = np.where(net <= 0, 0.0, gen_price)
    price = np.where((x_d > 0) & (x_d < q_d), a_d, np.where((x_c > 0) & (x_c < q_c), a_c, gen_price))
    return x_d, x_c, price


class MarketOracle:
    """``load``: [D, 288 + k] MW; ``moer``: [D, 288 + k] kg CO2 / MWh."""

    def __init__(self, params: MarketParams, load: np.ndarray, moer: np.ndarray, num_envs: int):
        self.p, self.load, self.moer, self.B = params, np.asarray(load, float), np.asarray(moer, float), num_envs

    def reset(self, day):
        B, k = self.B, self.p.k
        self.day = np.asarray(day, np.int64).copy()
        self.t = np.zeros(B, np.int64)
        self.e = np.full(B, self.p.e_init)
        self.prev_a = np.zeros((B, 2 * k))
        self.x_prev = np.zeros(B); self.p_prev = np.zeros(B)
        self.acc = np.zeros(B)
        return self.obs()

    def obs(self):
        k, t, d = self.p.k, self.t, self.day
        l_prev = np.where(t > 0, self.load[d, np.maximum(t - 1, 0)], 0.0)
        m_prev = np.where(t > 0, self.moer[d, np.maximum(t - 1, 0)], 0.0)
        lhat = self.load[d[:, None], t[:, None] + np.arange(k)[None]]
        mhat = self.moer[d[:, None], t[:, None] + np.arange(k)[None]]
        return np.concatenate([t[:, None].astype(float), self.e[:, None], self.prev_a, self.x_prev[:, None],
                               self.p_prev[:, None], l_prev[:, None], lhat, m_prev[:, None], mhat], 1).astype(np.float32)

    def step(self, action):
        p, k = self.p, self.p.k
        a = np.maximum(np.asarray(action, np.float32).astype(np.float64), 0.0)      # bids are non-negative
        a_c, a_d = a[:, 0], a[:, k]
        load, moer = self.load[self.day, self.t], self.moer[self.day, self.t]
        x_d, x_c, price = clear(p, load, self.e, a_c, a_d)
        x = (x_d - x_c) * TAU                                                       # MWh sold (+) / bought (-)
        self.e = np.clip(self.e - x_d * TAU / p.eta_d + x_c * TAU * p.eta_c, 0.0, p.e_max)
        r = price * x + p.p_co2 * moer * x
        self.t = self.t + 1
        term = self.t >= T_STEPS
        r = r - np.where(term, p.terminal_price * np.maximum(0.0, p.e_init - self.e), 0.0)
        if p.use_intermediate_rewards:
            reward = r
        else:
            self.acc = self.acc + r
            reward = np.where(term, self.acc, 0.0)
        self.prev_a, self.x_prev, self.p_prev = a.copy(), x, price
        return self.obs(), reward, term, np.zeros(self.B, bool), {"price": price, "dispatch": x, "energy": self.e.copy()}
