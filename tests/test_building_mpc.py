"""CPU: the BuildingEnv MPC agent (sustaingym/algorithms/building/mpc_controller.py:9-116 restated on HiGHS) — optimality
certificates of its plans and closed-loop behaviour on the CPU oracle env."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import building as OB
from sustaingym_b200.algorithms.building.mpc_controller import MPCAgent
from sustaingym_b200.envs.building.parameters import ParameterGenerator


class OracleEnv:
    """The attributes MPCAgent reads from a BuildingEnv, on top of the NumPy oracle (one env)."""

    def __init__(self, params: dict, epoch: int, T0=None):
        self.p = OB.BuildingParams.from_parameters(params)
        self.orc = OB.BuildingOracle(self.p, 1)
        self.n = self.p.n
        self.A_d, self.B_d = self.p.A_d, self.p.BD_d
        self.target, self.ac_map = self.p.target, self.p.ac_map
        self.out_temp, self.ground_temp, self.ghi, self.metabolism = (self.p.out_temp, self.p.ground_temp, self.p.ghi,
                                                                     self.p.metabolism)
        self.is_continuous_action = True
        self.Qlow = -self.ac_map
        self.action_space = SimpleNamespace(shape=(self.n,))
        self.state = self.orc.reset(np.array([epoch]), T0)[0]

    @property
    def epoch(self):
        return int(self.orc.epoch[0])

    def step(self, a):
        s, r, d, _ = self.orc.step(np.asarray(a, np.float32)[None, :])
        self.state = s[0]
        return self.state, float(r[0]), bool(d[0])


@pytest.fixture(scope="module")
def params():
    return ParameterGenerator("OfficeSmall", "Hot_Dry", "Tucson")


def plan_cost(env, ag, U):
    """The program's objective for the action plan U [P, n], evaluated independently of the agent's LP."""
    x = np.asarray(env.state[:env.n], np.float64)
    drift = (env.B_d[:, 2] * ag.temp + env.B_d[:, 1] * ag.ground_temp + env.B_d[:, 0] * ag.Occupower + env.B_d[:, -1] * ag.ghi)
    tot = 0.0
    for t in range(U.shape[0]):
        tot += ag.beta * np.linalg.norm(env.ac_map * (x - env.target), ag.pnorm) + (1 - ag.beta) * 24 * np.linalg.norm(U[t], ag.pnorm)
        x = env.A_d @ x + env.B_d[:, 3:-1] @ U[t] + drift
    return tot


@pytest.mark.parametrize("pnorm", [1, 2])
def test_plans_are_optimal(params, pnorm):
    env = OracleEnv(params, epoch=1500, T0=np.array([24.5, 23.0, 25.0, 26.0, 22.5, 24.0]))
    rng = np.random.default_rng(0)
    for P in (1, 2, 5):
        ag = MPCAgent(env, beta=0.9, pnorm=pnorm, planning_steps=P)
        action, state = ag.predict()
        assert action.shape == (env.n,) and np.all(np.abs(action) <= 1 + 1e-9)
        if P == 1:           # the reference's program does not price the state the only action leads to
            assert np.allclose(action, 0, atol=1e-9)
            continue
        assert ag.gap <= 1e-7 * max(1.0, ag.objective)
        # recover the whole plan from the agent's solution by re-solving for it is not possible from outside: compare
        # the certified objective with the independent evaluation of perturbed plans around the returned first action
        drift = (env.B_d[:, 2] * ag.temp + env.B_d[:, 1] * ag.ground_temp + env.B_d[:, 0] * ag.Occupower + env.B_d[:, -1] * ag.ghi)
        assert np.allclose(state, env.A_d @ env.state[:env.n].astype(np.float64) + env.B_d[:, 3:-1] @ action + drift, atol=1e-9)
        best = ag.objective
        for _ in range(300):  # no plan does better than the certified optimum (random plans + plans near the first action)
            U = np.clip(np.tile(action, (P, 1)) + rng.normal(0, rng.choice([1e-3, 1e-2, 0.1, 0.5]), (P, env.n)), -1, 1)
            assert plan_cost(env, ag, U) >= best - 1e-6 * max(1.0, best)


def test_mpc_model_is_the_env_model_and_beats_doing_nothing(params):
    """Closed loop on the oracle env: the state the agent predicts for its action is the state the env reaches (f32
    rounding of the env's state apart), and over 36 steps from a warm building the MPC earns more than zero action."""
    T0 = np.array([25.5, 24.0, 26.0, 27.0, 23.5, 25.0])
    env = OracleEnv(params, epoch=1439, T0=T0)
    ag = MPCAgent(env, beta=float(params["reward_beta"]), pnorm=2, planning_steps=4, gap_tol=1e-6)
    total = 0.0
    for _ in range(36):
        a, pred = ag.predict()
        s, r, _ = env.step(a)
        np.testing.assert_allclose(s[:env.n], pred, atol=2e-4)
        total += r
    idle = OracleEnv(params, epoch=1439, T0=T0)
    total_idle = sum(idle.step(np.zeros(idle.n))[1] for _ in range(36))
    assert total > total_idle
    assert np.abs(env.state[:env.n] - env.target).mean() < np.abs(idle.state[:idle.n] - idle.target).mean()
