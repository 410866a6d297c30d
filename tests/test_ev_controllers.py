"""CPU: the optimisation-based EV controllers (sustaingym/algorithms/evcharging/baselines.py:53-228 restated as LPs with
cutting planes for the magnitude cones) — the solver against an independent nonlinear solve, MPC against closed-form
answers, and both controllers played on the CPU oracle env of a whole day."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import ev as O
from sustaingym_b200.algorithms.evcharging.optimal import MPC, MagnitudeCuts, OfflineOptimal, _solve_lp_with_cuts
from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
from sustaingym_b200.envs.evcharging.network import caltech_acn

from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv as _E

# env.py:85-99 of the reference, as the engine states them
A_PERS_TO_KWH, PROFIT, CARBON = float(_E.A_PERS_TO_KWH), float(_E.PROFIT_FACTOR), float(_E.CARBON_COST_FACTOR)


def stub_env(cn, moer=None, evs=()):
    """What the controllers read from an EVChargingEnv, without the CUDA engine behind it."""
    consts = SimpleNamespace(ACTION_SCALE_FACTOR=32, PROFIT_FACTOR=PROFIT, CARBON_COST_FACTOR=CARBON,
                             A_PERS_TO_KWH=A_PERS_TO_KWH)
    return SimpleNamespace(action_space=SimpleNamespace(low=np.zeros(54)), moer_forecast_steps=36, _b=consts, cn=cn,
                           num_stations=cn.num_stations, moer=moer, _evs=list(evs))


def test_constants():
    assert A_PERS_TO_KWH == pytest.approx(0.208 / 12) and PROFIT == pytest.approx(A_PERS_TO_KWH * 0.03)
    assert CARBON == pytest.approx(A_PERS_TO_KWH * 30.85 / 1000)


def test_cutting_plane_lp_is_bracketed_by_a_fine_polyhedral_solve():
    """48 stations want to charge during 2 periods: pods AND transformer windings bind.  Independent check: the same
    program with every cone replaced by K = 720 fixed half-planes (no cut generation) is a relaxation whose optimum U
    is an upper bound, and its solution shrunk by cos(pi / K) is feasible for the cones, i.e. a lower bound:
    cos(pi / K) * U <= optimum <= U."""
    from scipy.optimize import linprog
    cn = caltech_acn()
    cuts = MagnitudeCuts(cn, 32.0)
    rng = np.random.default_rng(0)
    stations = np.sort(rng.choice(54, 48, replace=False))
    n, L = stations.size, 2
    var = np.arange(n * L).reshape(n, L)
    cost = (rng.uniform(0.5, 1.0, L)[None, :] * np.ones((n, 1))).ravel()
    rows = np.zeros((n, n * L))
    for i in range(n):
        rows[i, var[i]] = 1
    budget = rng.uniform(0.5, 2.0, n)
    x = _solve_lp_with_cuts(cost, np.ones(n * L), rows, budget, [(var[:, t], stations) for t in range(L)], cuts)
    X = np.zeros((54, L)); X[stations] = x.reshape(n, L)
    assert cuts.max_violation(X) <= 2e-8 and (rows @ x <= budget + 1e-9).all() and x.min() >= -1e-12 and x.max() <= 1 + 1e-12
    z = np.abs(cuts.At @ X)
    assert (z[:6] > 0.999 * cuts.r[:6, None]).any() and (z[6:] > 0.999 * cuts.r[6:, None]).any()   # both kinds active
    K = 720
    At = cuts.At[:, stations]
    A, b = [rows], [budget]
    for t in range(L):
        for j in range(At.shape[0]):
            blk = np.zeros((K, n * L))
            blk[:, var[:, t]] = (np.exp(-1j * 2 * np.pi * np.arange(K) / K)[:, None] * At[j][None, :]).real
            A.append(blk); b.append(np.full(K, cuts.r[j]))
    ref = linprog(-cost, A_ub=np.vstack(A), b_ub=np.concatenate(b), bounds=[(0, 1)] * (n * L), method="highs")
    assert ref.status == 0
    U = -ref.fun
    assert np.cos(np.pi / K) * U - 1e-9 <= cost @ x <= U + 1e-9


def test_mpc_closed_forms():
    cn = caltech_acn()
    mpc = MPC(stub_env(cn), lookahead=12)
    obs = {"demands": np.zeros(54, np.float32), "est_departures": np.zeros(54, np.float32),
           "forecasted_moer": np.linspace(0.2, 0.6, 36).astype(np.float32), "prev_moer": np.float32([0.2]),
           "timestep": np.float32([0.1])}
    assert np.array_equal(mpc.get_action(obs), np.zeros(54))                     # empty garage
    # three EVs far apart in the network, MOER rising: charge as early as possible, 1.0 per period until the budget
    # (demand / A_PERS_TO_KWH / 32 pilots) is spent or the estimated departure reached
    for i, (kwh, est) in {0: (1.0, 5), 20: (0.3, 9), 40: (5.0, 2)}.items():
        obs["demands"][i], obs["est_departures"][i] = kwh, est
    a = mpc.get_action(obs)
    units = {i: float(obs["demands"][i]) / A_PERS_TO_KWH / 32 for i in (0, 20, 40)}
    for i, est in ((0, 5), (20, 9), (40, 2)):
        want = np.zeros(12)
        left = units[i]
        for t in range(est):
            want[t] = min(1.0, left); left -= want[t]
        np.testing.assert_allclose(mpc.traj[i], want, atol=1e-9)
    assert a[0] == pytest.approx(1.0) and a[20] == pytest.approx(min(1.0, units[20])) and a[40] == pytest.approx(1.0)
    assert np.count_nonzero(a) == 3
    # MOER falling steeply enough that waiting pays: the budget goes to the LAST periods before the departure
    obs["forecasted_moer"] = np.linspace(0.9, 0.1, 36).astype(np.float32)
    mpc.get_action(obs)
    assert mpc.traj[20, 8] == pytest.approx(min(1.0, units[20])) and mpc.traj[20, :8].sum() == pytest.approx(max(0.0, units[20] - 1.0))
    # a departure estimate already in the past counts as one more period (baselines.py:121-123)
    obs["est_departures"][0] = -3
    mpc.get_action(obs)
    assert mpc.traj[0, 0] == pytest.approx(1.0) and mpc.traj[0, 1:].sum() == 0
    # a full pod: the sum of its pilots meets the pod limit 80 A / 32 at every period, nothing else is touched
    cls, _ = cn.class_reduce()
    pod = np.nonzero(cn.constraint_matrix[6] != 0)[0]
    obs["demands"][:] = 0; obs["est_departures"][:] = 0
    obs["demands"][pod] = 20.0; obs["est_departures"][pod] = 30
    a = mpc.get_action(obs)
    assert a[pod].sum() == pytest.approx(80 / 32, rel=1e-9) and np.count_nonzero(np.delete(a, pod)) == 0
    assert mpc._cuts.max_violation(mpc.traj) <= 2e-8


def play(orc, gen, seed, controller, env_stub):
    days = gen.first_days(np.asarray([seed]), 1)
    obs = orc.reset(days, gen.moer_days(days))
    env_stub.moer = gen.moer[gen.moer_days(days)[0]].astype(np.float64)
    ids = env_stub.cn.station_ids
    env_stub._evs = [SimpleNamespace(station_id=ids[st], arrival=a, departure=d, estimated_departure=e, requested_energy=k)
                     for a, d, e, k, st in gen.days[days[0]]]
    controller.reset()
    total, worst = 0.0, 0.0
    for _ in range(288):
        a = np.asarray(controller.get_action({k: v[0] for k, v in obs.items()}), dtype=np.float32)
        obs, r, d, _, info = orc.step(a[None, :])
        total += float(r[0])
    return total, info


def test_controllers_on_a_whole_day_of_the_oracle_env():
    """OfflineOptimal >= MPC >= greedy-with-projection on a busy synthetic day (return = profit - carbon - violation);
    the optimal trajectory never needs the env's projection to be feasible."""
    gen = SyntheticTraceGenerator("caltech", num_days=2, mean_sessions=90.0, seed=5)
    net = O.caltech_acn()
    cn = caltech_acn()
    orc = O.EVOracle(net, [[O.Session(*x) for x in d] for d in gen.days], gen.moer.astype(np.float64), 1, 36, True)
    env = stub_env(cn)
    greedy = SimpleNamespace(get_action=lambda o: np.where(o["demands"] > 0, 1.0, 0.0), reset=lambda: None)
    r_greedy, _ = play(orc, gen, 0, greedy, env)
    r_mpc, _ = play(orc, gen, 0, MPC(env, lookahead=12), env)
    off = OfflineOptimal(env)
    r_off, _ = play(orc, gen, 0, off, env)
    assert off.traj.shape == (54, 288) and off._cuts.max_violation(off.traj) <= 2e-8
    assert r_off > 0 and r_mpc > 0
    assert r_off >= r_mpc * (1 - 2e-3) and r_mpc >= r_greedy * (1 - 2e-3), (r_greedy, r_mpc, r_off)
    # the LP's objective bounds what any controller can earn in the env from above (pilot rounding, battery tapering)
    w = 32 * (PROFIT - CARBON * env.moer[1:289, 0])
    assert (off.traj * w[None, :]).sum() >= r_off - 1e-9
