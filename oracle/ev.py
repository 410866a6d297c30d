"""TEST INFRASTRUCTURE — CPU oracle for the EVChargingEnv hot path.  PARITY UNPINNED (see below).

What it restates (reference = chrisyeh96/sustaingym @ eac2a4d, file:line under /root/reference):

* glue that IS in the reference tree
    - constants                          sustaingym/envs/evcharging/env.py:99-114
    - ``to_schedule``   (a*32, rounding) env.py:340-379
    - ``observation``                    env.py:381-394
    - ``reward``                         env.py:431-464
    - ``project_action`` problem data    env.py:178-221, 473-500  (solver: cvxpy+MOSEK,
                                         sustaingym/envs/utils.py:6-24)
    - event-queue construction           sustaingym/envs/evcharging/event_generation.py:149-207
    - DiscreteActionWrapper              sustaingym/envs/wrappers.py:43-45
* arithmetic that lives in the UN-VENDORED third-party package ``acnportal`` (pyproject.toml:28
  pins ``acnportal>=0.3.3``; not installed here, no network).  Its published algorithm is
  restated from the acnportal 0.3.x sources as documented in SURVEY.md Appendix A:
    - ``caltech_acn``                    acnportal/acnsim/network/sites/caltech_acn.py
    - ``Simulator.step`` / events        acnportal/acnsim/simulator.py, events/event.py, event_queue.py
    - ``ChargingNetwork.constraint_current`` acnportal/acnsim/network/network.py
    - ``FiniteRatesEVSE``, ``EV.charge``, ``Linear2StageBattery._charge_stepwise``
                                         acnportal/acnsim/models/{evse,ev,battery}.py
    - ``Interface.active_sessions``      acnportal/acnsim/interface.py

PARITY UNPINNED: neither acnportal, cvxpy nor MOSEK can be imported in the build container and
the reference's own tests hold no numeric vectors for this path, so this oracle
is pinned only by (i) the reference's glue code it transcribes, (ii) the station order verified
against the reference's GMM pickles, (iii) structural identities
(``tests/test_oracle_ev.py``) and (iv) for the projection, an optimality certificate (KKT
residuals) plus an independent SciPy SLSQP solve.  MOSEK's own answer is only accurate to its
tolerances (~1e-8), so even the real reference is not bit-reproducible on this sub-step.

Two implementations live here on purpose:
  ``ScalarEVSim``   one env, Python objects and a heap-ordered event queue — a literal
                    transcription of the acnsim control flow (slow, used to pin the next one);
  ``EVOracle``      B envs vectorised with NumPy (used as the checker for the CUDA path and as
                    the CPU baseline in bench.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field

import numpy as np

# ---- env.py:99-114 -----------------------------------------------------------------------------
TIMESTEP_DURATION = 5
ACTION_SCALE_FACTOR = 32
VOLTAGE = 208
MARGINAL_REVENUE_PER_KWH = 0.15
OPERATING_MARGIN = 0.20
MARGINAL_PROFIT_PER_KWH = MARGINAL_REVENUE_PER_KWH * OPERATING_MARGIN
CO2_COST_PER_METRIC_TON = 30.85
A_MINS_TO_KWH = (1 / 60) * (VOLTAGE / 1000)
VIOLATION_WEIGHT = 0.001
A_PERS_TO_KWH = A_MINS_TO_KWH * TIMESTEP_DURATION
PROFIT_FACTOR = A_PERS_TO_KWH * MARGINAL_PROFIT_PER_KWH
VIOLATION_FACTOR = A_PERS_TO_KWH * VIOLATION_WEIGHT
CARBON_COST_FACTOR = A_PERS_TO_KWH * (CO2_COST_PER_METRIC_TON / 1000)
MAX_TIMESTEP = 288
# event_generation.py:60-62
BATTERY_CAPACITY = 100.0
MAX_POWER = 100.0
TRANSITION_SOC = 0.8          # acnportal Linear2StageBattery default
FULLY_CHARGED_EPS = 1e-3      # acnportal EV.fully_charged: not (remaining_demand > 1e-3)


# =================================================================================================
# Network (acnportal caltech_acn, defaults basic_evse=False, voltage=208, transformer_cap=150)
# =================================================================================================
@dataclass
class Network:
    station_ids: list
    phase_deg: np.ndarray          # [n]
    evse_type: list                # 'AV' -> {0,6..32} A ; 'CC' -> {0,8,16,24,32} A
    constraint_matrix: np.ndarray  # [J, n] real
    magnitudes: np.ndarray         # [J] A
    constraint_names: list = field(default_factory=list)

    @property
    def n(self) -> int:
        return len(self.station_ids)

    @property
    def min_pilot(self) -> np.ndarray:
        # AeroVironment allowable rates [0] + 6..32 -> min non-zero 6; ClipperCreek [0,8,..] -> 8
        return np.array([6 if t == "AV" else 8 for t in self.evse_type])

    @property
    def A_tilde(self) -> np.ndarray:
        """env.py:485-486: constraint_matrix * exp(1j*deg2rad(phase))."""
        return self.constraint_matrix * np.exp(1j * np.deg2rad(self.phase_deg))[None, :]

    def constraint_current(self, schedule: np.ndarray) -> np.ndarray:
        """acnportal ChargingNetwork.constraint_current (non-linear form). schedule [..., n] A."""
        return schedule @ self.A_tilde.T


class _Current:
    """acnportal ``Current``: a sparse {station: coefficient} row supporting + - and scalar *."""

    def __init__(self, ids=None, coeffs=None):
        self.c = dict(coeffs) if coeffs is not None else {i: 1.0 for i in (ids or [])}

    def __add__(self, o):
        out = dict(self.c)
        for k, v in o.c.items():
            out[k] = out.get(k, 0.0) + v
        return _Current(coeffs=out)

    def __sub__(self, o):
        return self + o * (-1.0)

    def __mul__(self, s):
        return _Current(coeffs={k: v * s for k, v in self.c.items()})

    __rmul__ = __mul__


def caltech_acn(voltage: float = 208, transformer_cap: float = 150) -> Network:
    cc_pod = ["CA-322", "CA-493", "CA-496", "CA-320", "CA-495", "CA-321", "CA-323", "CA-494"]
    av_pod = ["CA-324", "CA-325", "CA-326", "CA-327", "CA-489", "CA-490", "CA-491", "CA-492"]
    ab = [f"CA-{i}" for i in [308, 508, 303, 513, 310, 506, 316, 500, 318, 498]] + av_pod + cc_pod
    bc = [f"CA-{i}" for i in [304, 512, 305, 511, 313, 503, 311, 505, 317, 499, 148, 149, 212, 213]]
    ca = [f"CA-{i}" for i in [307, 509, 309, 507, 306, 510, 315, 501, 319, 497, 312, 504, 314, 502]]
    ids, phase, typ = [], [], []
    for e in ab:
        ids.append(e); phase.append(30.0); typ.append("CC" if e in cc_pod else "AV")
    for e in bc:
        ids.append(e); phase.append(-90.0); typ.append("AV")
    for e in ca:
        ids.append(e); phase.append(150.0); typ.append("AV")
    I3a = _Current(ab) - _Current(ca)
    I3b = _Current(bc) - _Current(ab)
    I3c = _Current(ca) - _Current(bc)
    I2a = (1 / 4) * (I3a - I3c)
    I2b = (1 / 4) * (I3b - I3a)
    I2c = (1 / 4) * (I3c - I3b)
    primary = transformer_cap * 1000 / 3 / 277
    secondary = transformer_cap * 1000 / 3 / 120
    rows = [("Secondary A", I3a, secondary), ("Secondary B", I3b, secondary), ("Secondary C", I3c, secondary),
            ("Primary A", I2a, primary), ("Primary B", I2b, primary), ("Primary C", I2c, primary),
            ("CC Pod", _Current(cc_pod), 80.0), ("AV Pod", _Current(av_pod), 80.0)]
    idx = {e: i for i, e in enumerate(ids)}
    M = np.zeros((len(rows), len(ids)))
    for r, (_, cur, _) in enumerate(rows):
        for k, v in cur.c.items():
            M[r, idx[k]] = v
    return Network(ids, np.array(phase), typ, M, np.array([r[2] for r in rows], float), [r[0] for r in rows])


# =================================================================================================
# Sessions / traces
# =================================================================================================
@dataclass
class Session:
    """One row of the events DataFrame (event_generation.py:125-147) after the cap at :169-170."""
    arrival: int
    departure: int
    est_departure: int
    energy: float        # requested kWh, already min(.., requested_energy_cap)
    station: int         # index into Network.station_ids


# =================================================================================================
# ScalarEVSim — literal single-env transcription
# =================================================================================================
class _Battery:
    """acnportal Linear2StageBattery(capacity=100, init_charge, max_power=100), stepwise, no noise."""

    def __init__(self, init_charge: float):
        self.capacity = BATTERY_CAPACITY
        self.charge_kwh = init_charge
        self.max_power = MAX_POWER

    def charge(self, pilot: float, voltage: float, period: float) -> float:
        rate_to_full = (self.capacity - self.charge_kwh) / (period / 60)
        soc = self.charge_kwh / self.capacity
        if soc < TRANSITION_SOC:
            power = min([pilot * voltage / 1000, self.max_power, rate_to_full])
        else:
            power = min([pilot * voltage / 1000,
                         (1 - soc) / (1 - TRANSITION_SOC) * self.max_power, rate_to_full])
        self.charge_kwh += power * (period / 60)
        return power * 1000 / voltage


class _EV:
    def __init__(self, s: Session):
        self.s = s
        self.battery = _Battery(max(0.0, BATTERY_CAPACITY - s.energy))   # event_generation.py:173-176
        self.energy_delivered = 0.0
        self.current_charging_rate = 0.0

    def charge(self, pilot, voltage, period):
        rate = self.battery.charge(pilot, voltage, period)
        self.energy_delivered += rate * voltage / 1000 * (period / 60)
        self.current_charging_rate = rate
        return rate

    @property
    def remaining_demand(self):
        return self.s.energy - self.energy_delivered

    @property
    def fully_charged(self):
        return not (self.remaining_demand > FULLY_CHARGED_EPS)


_PREC = {"unplug": 0, "plugin": 10, "recompute": 20}   # acnportal event precedences


class ScalarEVSim:
    """EVChargingEnv for ONE env with acnsim's control flow spelled out (env.py:229-338)."""

    def __init__(self, net: Network, sessions: list, moer: np.ndarray, moer_forecast_steps: int = 36,
                 project_action_in_env: bool = True):
        self.net, self.sessions, self.moer = net, sessions, np.asarray(moer, np.float64)
        self.k = moer_forecast_steps
        self.project = project_action_in_env
        self.reset()

    def reset(self):
        n = self.net.n
        self.evse = [None] * n                      # EV plugged at each EVSE
        self.queue, self._cnt = [], 0
        arrivals = set()
        for s in self.sessions:                     # event_generation.py:161-196
            self._push(s.arrival, "plugin", _EV(s))
            arrivals.add(s.arrival)
        for ts in range(MAX_TIMESTEP + 1):          # event_generation.py:199-203
            if ts not in arrivals:
                self._push(ts, "recompute", None)
        self.iteration = 0
        self.t = 0
        self.breakdown = {"profit": 0.0, "carbon_cost": 0.0, "excess_charge": 0.0}
        self.demands = np.zeros(n, np.float32)
        self.est_departures = np.zeros(n, np.float32)
        return self._observation()

    def _push(self, ts, kind, ev):
        # heap of (timestamp, event); events compare by precedence (acnportal Event.__lt__);
        # the counter only breaks exact ties, which acnsim leaves to heapq's sift order.
        heapq.heappush(self.queue, (ts, _PREC[kind], self._cnt, kind, ev))
        self._cnt += 1

    def _observation(self):
        self.est_departures.fill(0); self.demands.fill(0)                 # env.py:383-384
        for i, ev in enumerate(self.evse):                                # Interface.active_sessions()
            if ev is not None and not ev.fully_charged:
                self.est_departures[i] = ev.s.est_departure - self.t      # env.py:387
                self.demands[i] = ev.remaining_demand                     # env.py:388
        return {
            "timestep": np.array([self.t / MAX_TIMESTEP], np.float32),
            "est_departures": self.est_departures.copy(), "demands": self.demands.copy(),
            "prev_moer": np.array([self.moer[self.t, 0]], np.float32),
            "forecasted_moer": self.moer[self.t, 1:self.k + 1].astype(np.float32),
        }

    def step(self, action: np.ndarray):
        net = self.net
        self.t += 1                                                       # env.py:279
        if self.project:
            x = project_action(net, np.asarray(action)[None], self.demands[None])[0][0]
        else:
            x = np.asarray(action, np.float64)
        schedule = to_schedule(net, x[None])[0]                           # env.py:366-378
        # ---- acnsim Simulator.step(schedule) with _resolve forced False afterwards ----------
        rates = np.zeros(net.n)
        for i, ev in enumerate(self.evse):                                # update_pilots
            if ev is not None:
                rates[i] = ev.charge(schedule[i], VOLTAGE, TIMESTEP_DURATION)
        self.iteration += 1
        while self.queue and self.queue[0][0] <= self.iteration:          # get_current_events
            ts, _, _, kind, ev = heapq.heappop(self.queue)
            if kind == "plugin":
                if self.evse[ev.s.station] is not None:
                    raise RuntimeError("StationOccupiedError")
                self.evse[ev.s.station] = ev
                self._push(ev.s.departure, "unplug", ev)
            elif kind == "unplug":
                if self.evse[ev.s.station] is ev:
                    self.evse[ev.s.station] = None
        done = len(self.queue) == 0
        obs = self._observation()
        # ---- reward: env.py:444-462 ----------------------------------------------------------
        total = rates.sum()
        profit = PROFIT_FACTOR * total
        mag = np.abs(net.constraint_current(schedule))
        excess = np.sum(np.maximum(0, mag - net.magnitudes)) * VIOLATION_FACTOR
        carbon = CARBON_COST_FACTOR * total * self.moer[self.t, 0]
        reward = profit - carbon - excess
        self.breakdown["profit"] += profit
        self.breakdown["carbon_cost"] += carbon
        self.breakdown["excess_charge"] += excess
        return obs, reward, done, False, {"schedule": schedule, "rates": rates}


def flatten_obs(obs: dict) -> np.ndarray:
    """gymnasium 0.28 ``spaces.flatten`` of the Dict obs: keys in sorted order."""
    return np.concatenate([np.asarray(obs[k], np.float32).reshape(obs[k].shape[:-1] + (-1,))
                           for k in sorted(obs)], axis=-1)


# =================================================================================================
# shared glue (vectorised over a leading batch axis)
# =================================================================================================
def discrete_to_continuous(action_int: np.ndarray, bins: int) -> np.ndarray:
    """wrappers.py:43-45."""
    return np.asarray(action_int, dtype=np.float32) / (bins - 1)


def to_schedule(net: Network, x: np.ndarray) -> np.ndarray:
    """env.py:366-378. x: [B, n] normalised action (f32 or f64) -> pilots in A (f64 integers)."""
    a = np.asarray(x) * ACTION_SCALE_FACTOR
    is_av = (net.min_pilot == 6)[None, :]
    av = np.where(a >= 6, np.round(a), 0.0)
    cc = np.round(a / 8) * 8
    return np.where(is_av, av, cc).astype(np.float64)


# =================================================================================================
# action projection (env.py:178-221): argmin ||x-a||_2, 0<=x<=u, 32|A~x| <= magnitudes
# =================================================================================================
def projection_bounds(demands_f32: np.ndarray) -> np.ndarray:
    """env.py:188-189: min(1, demands / A_PERS_TO_KWH / 32) in f64."""
    d = np.asarray(demands_f32, np.float32).astype(np.float64)
    return np.minimum(1.0, d / A_PERS_TO_KWH / ACTION_SCALE_FACTOR)


def project_action(net: Network, action: np.ndarray, demands_f32: np.ndarray,
                   tol: float = 1e-13, max_iter: int = 100):
    """Returns (x [B,n] f64, info dict). Fast path: the box-clipped action when it already satisfies
    the network constraints (it is then the exact projection); otherwise a primal-dual
    interior-point solve of the QCQP (dense Newton systems — deliberately naive)."""
    a = np.asarray(action, np.float64)
    u = projection_bounds(demands_f32)
    x = np.clip(a, 0.0, u)
    At = net.A_tilde
    r = net.magnitudes / ACTION_SCALE_FACTOR
    viol = np.abs(x @ At.T) > r[None, :]
    slow = np.nonzero(viol.any(axis=1))[0]
    info = {"slow": slow, "iters": np.zeros(len(a), np.int32), "kkt": np.zeros(len(a))}
    if len(slow):
        xs, it, kkt = _ipm(a[slow], u[slow], At, r, tol, max_iter)
        x[slow] = xs
        info["iters"][slow] = it
        info["kkt"][slow] = kkt
    return x, info


def _ipm(a, u, At, r, tol, max_iter):
    """Batched primal-dual IPM.  a,u: [M,n]; At: [J,n] complex; r: [J] (normalised limits)."""
    M, n = a.shape
    J = len(r)
    Are, Aim = At.real, At.imag
    free = u > 0
    uf = np.where(free, u, 1.0)
    # strictly interior start: x = theta*u with every |A~x| <= r/2
    w_u = np.abs((u @ At.T))
    theta = np.minimum(0.5, 0.5 * np.min(r[None, :] / np.maximum(w_u, 1e-300), axis=1))
    x = np.where(free, theta[:, None] * uf, 0.0)
    mu = np.full(M, 0.1)
    zl = np.where(free, mu[:, None] / np.where(free, x, 1.0), 0.0)
    zu = np.where(free, mu[:, None] / np.where(free, uf - x, 1.0), 0.0)
    w = x @ At.T
    q = r[None, :] ** 2 - np.abs(w) ** 2
    lam = mu[:, None] / q
    iters = np.zeros(M, np.int32)
    active = np.ones(M, bool)
    tau = 0.995
    ncomp = 2 * free.sum(axis=1) + J
    for it in range(max_iter):
        w = x @ At.T
        q = r[None, :] ** 2 - np.abs(w) ** 2
        g = w.real[:, :, None] * Are[None] + w.imag[:, :, None] * Aim[None]          # [M,J,n] = -grad q / 2
        xs = np.where(free, x, 1.0)
        us = np.where(free, uf - x, 1.0)
        r_d = (x - a) - zl + zu + 2 * np.einsum("mj,mjn->mn", lam, g)
        r_d = np.where(free, r_d, 0.0)
        gap = (np.where(free, zl * xs + zu * us, 0.0).sum(axis=1) + (lam * q).sum(axis=1)) / ncomp
        resid = np.abs(r_d).max(axis=1)
        newly = active & (gap < tol) & (resid < 100 * tol)
        active &= ~newly
        if not active.any():
            break
        mu = 0.1 * gap
        D = np.where(free, 1.0 + zl / xs + zu / us, 1.0)
        H = np.einsum("mj,jn,jk->mnk", 2 * lam, Are, Are) + np.einsum("mj,jn,jk->mnk", 2 * lam, Aim, Aim) \
            + np.einsum("mj,mjn,mjk->mnk", 4 * lam / q, g, g)
        fm = free[:, :, None] & free[:, None, :]
        H = np.where(fm, H, 0.0)
        H[:, np.arange(n), np.arange(n)] += D
        rho = -(x - a) + mu[:, None] / xs - mu[:, None] / us - 2 * np.einsum("mj,mjn->mn", mu[:, None] / q, g)
        rho = np.where(free, rho, 0.0)
        dx = np.linalg.solve(H, rho[:, :, None])[:, :, 0]
        dx = np.where(free, dx, 0.0)
        dzl = np.where(free, mu[:, None] / xs - zl - zl / xs * dx, 0.0)
        dzu = np.where(free, mu[:, None] / us - zu + zu / us * dx, 0.0)
        dlam = mu[:, None] / q - lam + 2 * lam / q * np.einsum("mjn,mn->mj", g, dx)
        dw = dx @ At.T
        # fraction-to-boundary step
        with np.errstate(divide="ignore", invalid="ignore"):
            a1 = np.where(free & (dx < 0), -tau * xs / dx, np.inf).min(axis=1)
            a2 = np.where(free & (dx > 0), tau * us / dx, np.inf).min(axis=1)
            b = (w.conj() * dw).real
            c2 = np.abs(dw) ** 2
            den = b + np.sqrt(b * b + c2 * tau * q)
            a3 = np.where(den > 0, tau * q / den, np.inf).min(axis=1)
            a4 = np.where(free & (dzl < 0), -tau * zl / dzl, np.inf).min(axis=1)
            a5 = np.where(free & (dzu < 0), -tau * zu / dzu, np.inf).min(axis=1)
            a6 = np.where(dlam < 0, -tau * lam / dlam, np.inf).min(axis=1)
        alpha = np.minimum(1.0, np.minimum.reduce([a1, a2, a3, a4, a5, a6]))
        alpha = np.where(active, alpha, 0.0)[:, None]
        x = x + alpha * dx
        zl = zl + alpha * dzl
        zu = zu + alpha * dzu
        lam = lam + alpha * dlam
        iters += active
    # certificate: KKT residual at the returned point
    w = x @ At.T
    q = r[None, :] ** 2 - np.abs(w) ** 2
    g = w.real[:, :, None] * Are[None] + w.imag[:, :, None] * Aim[None]
    r_d = np.where(free, (x - a) - zl + zu + 2 * np.einsum("mj,mjn->mn", lam, g), 0.0)
    comp = np.maximum(np.where(free, np.maximum(zl * x, zu * (uf - x)), 0.0).max(axis=1), (lam * q).max(axis=1))
    infeas = np.maximum(np.maximum(-q, 0).max(axis=1), np.maximum(np.maximum(-x, x - u), 0).max(axis=1))
    kkt = np.maximum.reduce([np.abs(r_d).max(axis=1), comp, infeas])
    return x, iters, kkt


# =================================================================================================
# EVOracle — B envs vectorised
# =================================================================================================
class EVOracle:
    """B independent EVChargingEnv instances in lock-step.

    ``days``: list (len D) of lists of ``Session``; ``moer``: [Dm, 289, 37] f64 (or f32) tables.
    Per-env session lists are flattened to padded [B, S] arrays at reset."""

    def __init__(self, net: Network, days: list, moer: np.ndarray, num_envs: int,
                 moer_forecast_steps: int = 36, project_action_in_env: bool = True):
        self.net, self.days, self.moer = net, days, np.asarray(moer, np.float64)
        self.B, self.k, self.project = num_envs, moer_forecast_steps, project_action_in_env
        self.n = net.n
        self.S = max(1, max(len(d) for d in days))
        D = len(days)
        self._arr = np.full((D, self.S), 10 ** 6, np.int64)
        self._dep = np.full((D, self.S), 10 ** 6, np.int64)
        self._est = np.zeros((D, self.S), np.int64)
        self._req = np.zeros((D, self.S), np.float64)
        self._stn = np.zeros((D, self.S), np.int64)
        for d, ss in enumerate(days):
            for j, s in enumerate(sorted(ss, key=lambda s: (s.arrival, s.station))):
                self._arr[d, j], self._dep[d, j], self._est[d, j] = s.arrival, s.departure, s.est_departure
                self._req[d, j], self._stn[d, j] = s.energy, s.station
        self.day = np.zeros(num_envs, np.int64)
        self.moer_day = np.zeros(num_envs, np.int64)
        self.reset(np.zeros(num_envs, np.int64), np.zeros(num_envs, np.int64))

    def reset(self, day: np.ndarray, moer_day: np.ndarray, mask: np.ndarray | None = None):
        B, S = self.B, self.S
        if mask is None:
            mask = np.ones(B, bool)
            self.t = np.zeros(B, np.int64)
            self.delivered = np.zeros((B, S))
            self.charge = np.zeros((B, S))
            self.plugged = np.zeros((B, S), bool)
            self.gone = np.zeros((B, S), bool)
            self.breakdown = np.zeros((B, 3))
            self.demands = np.zeros((B, self.n), np.float32)
        self.day[mask] = np.asarray(day)[mask]
        self.moer_day[mask] = np.asarray(moer_day)[mask]
        self.t[mask] = 0
        self.delivered[mask] = 0
        self.charge[mask] = np.maximum(0.0, BATTERY_CAPACITY - self._req[self.day[mask]])
        self.plugged[mask] = False
        self.gone[mask] = False
        self.breakdown[mask] = 0
        return self._observation()

    def _observation(self):
        B, n = self.B, self.n
        d = self.day
        rem = self._req[d] - self.delivered
        act = self.plugged & (rem > FULLY_CHARGED_EPS)
        dem = np.zeros((B, n), np.float32)
        est = np.zeros((B, n), np.float32)
        bi, si = np.nonzero(act)
        dem[bi, self._stn[d][bi, si]] = rem[bi, si]
        est[bi, self._stn[d][bi, si]] = self._est[d][bi, si] - self.t[bi]
        self.demands = dem
        row = self.moer[self.moer_day, self.t]           # [B, 37]
        return {
            "timestep": (self.t / MAX_TIMESTEP).astype(np.float32)[:, None],
            "est_departures": est, "demands": dem,
            "prev_moer": row[:, :1].astype(np.float32),
            "forecasted_moer": row[:, 1:self.k + 1].astype(np.float32),
        }

    def step(self, action: np.ndarray):
        net, B, n = self.net, self.B, self.n
        self.t = self.t + 1
        if self.project:
            x, pinfo = project_action(net, action, self.demands)
        else:
            x, pinfo = np.asarray(action, np.float64), {"slow": np.array([], int)}
        schedule = to_schedule(net, x)                                   # [B, n]
        d = self.day
        stn = self._stn[d]
        pilot = np.take_along_axis(schedule, stn, axis=1)                # [B, S]
        # ---- Linear2StageBattery._charge_stepwise for plugged EVs -------------------------------
        period_h = TIMESTEP_DURATION / 60
        rate_to_full = (BATTERY_CAPACITY - self.charge) / period_h
        soc = self.charge / BATTERY_CAPACITY
        p_pilot = pilot * VOLTAGE / 1000
        stage1 = np.minimum(np.minimum(p_pilot, MAX_POWER), rate_to_full)
        stage2 = np.minimum(np.minimum(p_pilot, (1 - soc) / (1 - TRANSITION_SOC) * MAX_POWER), rate_to_full)
        power = np.where(soc < TRANSITION_SOC, stage1, stage2)
        power = np.where(self.plugged, power, 0.0)
        self.charge = self.charge + power * period_h
        rate = power * 1000 / VOLTAGE
        self.delivered = self.delivered + rate * VOLTAGE / 1000 * period_h
        total = rate.sum(axis=1)
        # ---- events with timestamp <= t: unplug (prec 0) before plugin (prec 10) -----------------
        t = self.t[:, None]
        arr, dep = self._arr[d], self._dep[d]
        newly = (~self.plugged) & (~self.gone) & (arr <= t)
        self.plugged |= newly
        leave = self.plugged & (dep <= t)
        self.plugged &= ~leave
        self.gone |= leave
        done = self.t >= MAX_TIMESTEP
        obs = self._observation()
        # ---- reward --------------------------------------------------------------------------
        profit = PROFIT_FACTOR * total
        mag = np.abs(net.constraint_current(schedule))
        excess = np.maximum(0, mag - net.magnitudes[None]).sum(axis=1) * VIOLATION_FACTOR
        carbon = CARBON_COST_FACTOR * total * self.moer[self.moer_day, self.t, 0]
        reward = profit - carbon - excess
        self.breakdown += np.stack([profit, carbon, excess], 1)
        info = {"schedule": schedule, "x": x, "slow": pinfo["slow"], "total_rate": total}
        return obs, reward, done, np.zeros(B, bool), info
