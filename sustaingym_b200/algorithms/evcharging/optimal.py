"""Optimisation-based controllers for EVChargingEnv: receding-horizon ``MPC`` and the hindsight ``OfflineOptimal``.

The reference states both as cvxpy programs handed to MOSEK (sustaingym/algorithms/evcharging/baselines.py:53-133 and
:136-228): a linear objective (profit minus carbon cost of the normalised pilots), box / mask bounds, per-session
energy budgets and the network's magnitude constraints ``|A~ x_t| <= m / 32`` at every period.  Neither cvxpy nor MOSEK
exists in this image, so the same programs are solved here as linear programs (HiGHS, through SciPy) with the second-order
cones handled by OUTER APPROXIMATION: a constraint whose columns share one phase (the pods: a real sum) is one exact
linear row; for the others (transformer windings: complex sums of several phases) the supporting half-planes
``Re(e^{-i phi} A~_j x_t) <= r_j`` at the angles of the columns seed the relaxation, the LP is solved, and for every
(constraint, period) pair still violated the half-plane at the violating angle is added — Kelley's cutting planes, which
terminate here after a few rounds because few cones are ever active (the loop stops at a relative violation of 1e-8,
HiGHS running with feasibility tolerances of 1e-9).

Parity: unpinned against MOSEK.  The programs are linear, so their optimal VALUE is what both solvers agree on (to solver
tolerance); where the optimal set is not a single point (ties between periods of equal MOER) HiGHS returns a vertex and
an interior-point solver a point in the face's interior — the tests therefore check feasibility, optimality certificates
(against an independent solve) and the ordering OfflineOptimal >= MPC >= Greedy on whole days, not MOSEK's iterates.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from ..base import BaseAlgorithm

MAX_ACTION = 1


class MagnitudeCuts:
    """Outer approximation of ``|A~_j x| <= r_j`` (j = 0..J-1) for nonnegative x over a subset of stations.

    ``rows(cols)`` gives the seed rows for the stations ``cols`` of one period; ``violated(x, cols)`` returns the new
    rows the point ``x`` calls for (empty when it satisfies every constraint to the tolerance)."""

    def __init__(self, cn, scale: float, tol: float = 1e-8):
        self.At = cn.constraint_matrix * np.exp(1j * np.deg2rad(cn.phase_deg))[None, :]     # env.py:485-486
        self.r = np.asarray(cn.magnitudes, dtype=np.float64) / scale
        self.tol = tol

    def seed_rows(self, cols: np.ndarray) -> list[tuple[np.ndarray, float]]:
        out = []
        for j in range(self.At.shape[0]):
            a = self.At[j, cols]
            nz = np.abs(a) > 0
            if not nz.any():
                continue
            angles = np.unique(np.round(np.angle(a[nz]), 12))
            for phi in angles:                      # one angle: the constraint is that single (exact) row
                out.append(((np.exp(-1j * phi) * a).real, self.r[j]))
        return out

    def violated(self, x: np.ndarray, cols: np.ndarray) -> list[tuple[np.ndarray, float]]:
        z = self.At[:, cols] @ x
        out = []
        for j in np.nonzero(np.abs(z) > self.r * (1 + self.tol))[0]:
            out.append(((np.exp(-1j * np.angle(z[j])) * self.At[j, cols]).real, self.r[j]))
        return out

    def max_violation(self, x_full: np.ndarray) -> float:
        """Largest relative excess of ``|A~ x_t|`` over its limit for a full ``[n]`` or ``[n, T]`` trajectory."""
        z = np.abs(self.At @ x_full)
        return float(np.max(z / (self.r[:, None] if z.ndim == 2 else self.r)) - 1.0)


def _solve_lp_with_cuts(cost, ub, budget_rows, budget_rhs, period_cols, cuts: MagnitudeCuts, max_rounds: int = 60):
    """maximise cost @ x  s.t. 0 <= x <= ub, budget_rows @ x <= budget_rhs and, for every period, the magnitude
    constraints on that period's variables (``period_cols``: list of (variable indices, station indices))."""
    from scipy.optimize import linprog
    from scipy.sparse import coo_matrix, vstack, csr_matrix
    nv = cost.size
    if nv == 0:
        return np.zeros(0)
    ri, ci, vv, rhs = [], [], [], []

    def add(rows, var_idx):
        for coef, r in rows:
            k = len(rhs)
            nzc = np.nonzero(coef)[0]
            ri.extend([k] * nzc.size); ci.extend(var_idx[nzc].tolist()); vv.extend(coef[nzc].tolist()); rhs.append(r)

    for var_idx, stations in period_cols:
        add(cuts.seed_rows(stations), var_idx)
    base = csr_matrix(budget_rows) if budget_rows is not None and budget_rows.shape[0] else None
    for _ in range(max_rounds):
        A_net = coo_matrix((vv, (ri, ci)), shape=(len(rhs), nv)).tocsr() if rhs else None
        mats = [m for m in (base, A_net) if m is not None]
        b = np.concatenate([v for v in (budget_rhs if base is not None else None, np.asarray(rhs) if rhs else None)
                            if v is not None]) if mats else None
        res = linprog(-cost, A_ub=vstack(mats).tocsr() if mats else None, b_ub=b,
                      bounds=np.stack([np.zeros(nv), ub], axis=1), method="highs",
                      options={"primal_feasibility_tolerance": 1e-9, "dual_feasibility_tolerance": 1e-9})
        if res.status != 0:
            raise RuntimeError(f"LP solve failed: {res.message}")
        x = res.x
        n_before = len(rhs)
        for var_idx, stations in period_cols:
            add(cuts.violated(x[var_idx], stations), var_idx)
        if len(rhs) == n_before:
            return x
    raise RuntimeError("magnitude constraints: cutting planes did not converge")


class MPC(BaseAlgorithm):
    """Model predictive control over ``lookahead`` periods (baselines.py:53-133): maximise forecast profit minus carbon
    cost of the pilot trajectory; an EV can be charged until its estimated departure (at least one more period), never
    beyond its remaining demand; network constraints at every period; the first action of the trajectory is applied.
    Future arrivals are unknown to it."""

    def __init__(self, env, lookahead: int = 12):
        super().__init__(env)
        assert hasattr(env.action_space, "low"), "MPC only supports continuous action space"
        assert lookahead <= env.moer_forecast_steps, "MPC lookahead must be less than forecasted timesteps"
        self.lookahead = lookahead
        B = env._b
        self._scale = float(B.ACTION_SCALE_FACTOR)
        self._pf, self._cf, self._kwh = float(B.PROFIT_FACTOR), float(B.CARBON_COST_FACTOR), float(B.A_PERS_TO_KWH)
        self._cuts = MagnitudeCuts(env.cn, self._scale)
        self.traj = np.zeros((env.num_stations, lookahead))

    def get_action(self, observation: dict[str, Any]) -> np.ndarray:
        n, L = self.env.num_stations, self.lookahead
        demands = np.asarray(observation["demands"], dtype=np.float64)
        moers = np.asarray(observation["forecasted_moer"], dtype=np.float64)[:L]
        # an estimated departure already passed: the car is assumed to stay one more period (baselines.py:121-123)
        est = np.maximum(1.0, np.asarray(observation["est_departures"])).astype(np.int32)
        est = np.where(demands > 0, est, 0)
        act = np.nonzero(est > 0)[0]
        self.traj = np.zeros((n, L))
        if act.size == 0:
            return self.traj[:, 0]
        w = self._scale * (self._pf - self._cf * moers)                        # objective weight of period t
        var = -np.ones((n, L), dtype=np.int64)
        for i in act:
            var[i, :min(int(est[i]), L)] = 1
        idx = np.nonzero(var.ravel() > 0)[0]
        var.ravel()[idx] = np.arange(idx.size)
        st, tt = np.divmod(idx, L)
        cost = w[tt]
        rows = np.zeros((act.size, idx.size))
        for k, i in enumerate(act):
            rows[k, var[i][var[i] >= 0]] = 1.0
        rhs = demands[act] / self._kwh / self._scale
        period_cols = [(var[:, t][var[:, t] >= 0], np.nonzero(var[:, t] >= 0)[0]) for t in range(L)
                       if (var[:, t] >= 0).any()]
        x = _solve_lp_with_cuts(cost, np.full(idx.size, float(MAX_ACTION)), rows, rhs, period_cols, self._cuts)
        self.traj[st, tt] = x
        return self.traj[:, 0]


class OfflineOptimal(BaseAlgorithm):
    """The controller that knows the day (baselines.py:136-228): on the first call after ``env.reset`` one program over
    all 288 periods with the true MOER values, true arrivals / departures and requested energies; afterwards the stored
    trajectory is played back.  As in the reference the remaining-demand chain of a session runs over
    ``[arrival, departure - 1)`` (baselines.py:198-207): its last period is bounded by the pilot limit only."""

    TOTAL_TIMESTEPS = 288

    def __init__(self, env):
        super().__init__(env)
        assert hasattr(env.action_space, "low"), "Offline optimal only supports continuous action space"
        B = env._b
        self._scale = float(B.ACTION_SCALE_FACTOR)
        self._pf, self._cf, self._kwh = float(B.PROFIT_FACTOR), float(B.CARBON_COST_FACTOR), float(B.A_PERS_TO_KWH)
        self._cuts = MagnitudeCuts(env.cn, self._scale)
        self.reset()

    def reset(self) -> None:
        self.timestep = 0
        self.traj = None

    def _solve(self) -> np.ndarray:
        env, T, n = self.env, self.TOTAL_TIMESTEPS, self.env.num_stations
        moers = env.moer[1:T + 1, 0]
        w = self._scale * (self._pf - self._cf * moers)
        station_idx = {evse: i for i, evse in enumerate(env.cn.station_ids)}
        var = -np.ones((n, T), dtype=np.int64)
        sessions = []
        for ev in env._evs:
            i, a, d = station_idx[ev.station_id], max(ev.arrival, 0), min(ev.departure, T)
            if a >= d:
                continue
            var[i, a:d] = 1
            sessions.append((i, a, d, ev.requested_energy / self._kwh / self._scale))
        idx = np.nonzero(var.ravel() > 0)[0]
        var.ravel()[idx] = np.arange(idx.size)
        st, tt = np.divmod(idx, T)
        from scipy.sparse import lil_matrix
        rows = lil_matrix((len(sessions), idx.size))
        rhs = np.zeros(len(sessions))
        for k, (i, a, d, dem) in enumerate(sessions):
            # demands[a] = dem, demands[t+1] = demands[t] - traj[t] >= 0 for t in [a, d-1)  <=>  sum traj[a:d-1] <= dem
            if d - 1 > a:
                rows[k, var[i, a:d - 1]] = 1.0
            rhs[k] = dem
        period_cols = [(var[:, t][var[:, t] >= 0], np.nonzero(var[:, t] >= 0)[0]) for t in range(T)
                       if (var[:, t] >= 0).any()]
        x = _solve_lp_with_cuts(w[tt], np.full(idx.size, float(MAX_ACTION)), rows.tocsr(), rhs, period_cols, self._cuts)
        traj = np.zeros((n, T))
        traj[st, tt] = x
        return traj

    def get_action(self, observation: dict[str, Any]) -> np.ndarray:
        if self.timestep == 0:
            self.traj = self._solve()
        if self.timestep < self.TOTAL_TIMESTEPS:
            action = self.traj[:, self.timestep]
        else:                                      # (the reference prints the IndexError and sends zeros)
            action = np.zeros(self.env.num_stations)
        self.timestep += 1
        return action


def mpc_policy(batched_env, lookahead: int = 12, workers: int | None = None):
    """MPC as a device policy for ``run_batched``: the batched observation goes to the host once per step, one linear
    program per env is solved there (a thread pool: HiGHS releases the GIL), the actions come back as one ``[B, n]`` CUDA
    tensor.  All envs still advance with ONE kernel launch per step; the controller, not the env, is the cost."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    from types import SimpleNamespace

    import torch
    consts = SimpleNamespace(ACTION_SCALE_FACTOR=batched_env.ACTION_SCALE_FACTOR, PROFIT_FACTOR=batched_env.PROFIT_FACTOR,
                             CARBON_COST_FACTOR=batched_env.CARBON_COST_FACTOR, A_PERS_TO_KWH=batched_env.A_PERS_TO_KWH)
    view = SimpleNamespace(action_space=SimpleNamespace(low=np.zeros(batched_env.num_stations)),
                           moer_forecast_steps=batched_env.moer_forecast_steps, _b=consts, cn=batched_env.cn,
                           num_stations=batched_env.num_stations)
    B = batched_env.num_envs
    ctrls = [MPC(view, lookahead) for _ in range(B)]
    pool = ThreadPoolExecutor(max_workers=workers or min(B, os.cpu_count() or 1))

    def policy(obs: dict):
        host = {k: v.detach().cpu().numpy() for k, v in obs.items()}
        acts = list(pool.map(lambda i: ctrls[i].get_action({k: host[k][i] for k in host}), range(B)))
        return torch.from_numpy(np.asarray(acts, dtype=np.float32)).to(batched_env.device)
    return policy
