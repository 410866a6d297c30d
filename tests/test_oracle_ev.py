"""CPU: pins for the EV oracle (no reference-run vectors exist for this env — SURVEY.md F6/F7).

  * the literal single-env transcription (ScalarEVSim) and the vectorised oracle agree;
  * structural identities the reference documents (288-step episodes, reward = profit - carbon -
    excess, zero action -> zero reward, timestep = t/288);
  * the projection satisfies its KKT certificate and matches an independent SciPy SLSQP solve;
  * the product's network description equals the oracle's acnportal-style construction.
"""
import numpy as np
import pytest

from oracle import ev as O
from sustaingym_b200.data.ev_traces import pack_sessions, synthetic_days, synthetic_moer
from sustaingym_b200.envs.evcharging.network import caltech_acn


@pytest.fixture(scope="module")
def net():
    return O.caltech_acn()


def test_network_spec_matches_oracle_construction(net):
    spec = caltech_acn()
    assert list(spec.station_ids) == net.station_ids and spec.num_stations == 54
    np.testing.assert_array_equal(spec.phase_deg, net.phase_deg)
    np.testing.assert_array_equal(spec.min_pilot_signals, net.min_pilot)
    key = lambda M, m: sorted((round(float(x), 9),) + tuple(np.round(r, 9)) for r, x in zip(M, m))  # noqa: E731
    assert key(spec.constraint_matrix, spec.magnitudes) == key(net.constraint_matrix, net.magnitudes)
    cls, coef = spec.class_reduce()
    assert coef.shape == (8, 5) and sorted(np.bincount(cls)) == [8, 8, 10, 14, 14]
    s = np.random.default_rng(0).uniform(0, 32, (16, 54))
    At = spec.constraint_matrix * np.exp(1j * np.deg2rad(spec.phase_deg))[None]
    cs = np.stack([s[:, cls == c].sum(axis=1) for c in range(5)], axis=1)
    np.testing.assert_allclose(s @ At.T, cs @ coef.T, atol=1e-9)
    # SURVEY A.4 closed forms
    S = {c: cs[:, c] for c in range(5)}
    ab = S[cls[0]] + S[cls[10]] + S[cls[18]]
    bc, ca = S[cls[26]], S[cls[40]]
    i3a = np.abs(s @ At.T)[:, list(spec.constraint_names).index("Secondary A")]
    np.testing.assert_allclose(i3a ** 2, ab ** 2 + ca ** 2 + ab * ca, rtol=1e-10)


def test_constants_match_reference_values():
    assert abs(O.A_PERS_TO_KWH - 0.017333333333333333) < 1e-15
    assert O.PROFIT_FACTOR == O.A_PERS_TO_KWH * (0.15 * 0.20)


@pytest.mark.parametrize("project", [False, True])
def test_scalar_and_vectorised_oracles_agree(net, project):
    days = synthetic_days(3, seed=1, mean_sessions=60)
    moer = synthetic_moer(2, seed=3).astype(np.float64)
    sess = [[O.Session(*x) for x in d] for d in days]
    B = 3
    orc = O.EVOracle(net, sess, moer, B, project_action_in_env=project)
    orc.reset(np.arange(B) % 3, np.arange(B) % 2)
    sims = [O.ScalarEVSim(net, sess[b % 3], moer[b % 2], project_action_in_env=project) for b in range(B)]
    rng = np.random.default_rng(5)
    tot = np.zeros(B)
    for t in range(288):
        a = rng.uniform(0, 1, (B, 54)).astype(np.float32)
        if t % 50 == 0:
            a[:] = 0
        obs, r, d, tr, info = orc.step(a)
        fo = O.flatten_obs(obs)
        tot += r
        for b in range(B):
            o2, r2, d2, _, i2 = sims[b].step(a[b])
            np.testing.assert_array_equal(i2["schedule"], info["schedule"][b])
            np.testing.assert_allclose(O.flatten_obs(o2), fo[b], rtol=1e-6, atol=1e-9)
            assert abs(r2 - r[b]) <= 1e-12 + 1e-9 * abs(r2)
            assert d2 == d[b] == (t == 287)
        assert np.allclose(obs["timestep"][:, 0], (t + 1) / 288)
        if t % 50 == 0 and not project:
            assert np.all(r == 0)                      # a == 0 -> reward == 0
    bd = orc.breakdown
    np.testing.assert_allclose(tot, bd[:, 0] - bd[:, 1] - bd[:, 2], rtol=1e-9)
    if project:
        assert np.all(bd[:, 2] < 1e-2)                 # projected actions (almost) never violate the network


def test_flatten_order_is_sorted_keys(net):
    days = synthetic_days(1, seed=0)
    moer = synthetic_moer(1, seed=0).astype(np.float64)
    sim = O.ScalarEVSim(net, [O.Session(*x) for x in days[0]], moer[0])
    obs = sim.reset()
    flat = O.flatten_obs(obs)
    assert flat.shape == (146,)
    assert flat[-1] == 0 and flat[-2] == np.float32(moer[0, 0, 0])
    np.testing.assert_array_equal(flat[108:144], moer[0, 0, 1:37].astype(np.float32))


def test_rounding_rules(net):
    x = np.zeros((1, 54))
    av, cc = 0, 18       # first AB station is AeroVironment, index 18 is the first ClipperCreek
    assert net.evse_type[av] == "AV" and net.evse_type[cc] == "CC"
    for amps, want_av, want_cc in [(5.99, 0, 8), (6.0, 6, 8), (6.5, 6, 8), (7.5, 8, 8), (3.9, 0, 0),
                                   (4.0, 0, 0), (4.01, 0, 8), (11.99, 12, 8), (12.0, 12, 16), (20.0, 20, 16),
                                   (31.5, 32, 32), (32.0, 32, 32)]:
        x[0, av] = x[0, cc] = amps / 32
        s = O.to_schedule(net, x)
        assert (s[0, av], s[0, cc]) == (want_av, want_cc), amps


def test_projection_certificate_and_slsqp(net):
    from scipy.optimize import minimize
    rng = np.random.default_rng(0)
    M = 120
    a = rng.uniform(0, 1, (M, 54)).astype(np.float32)
    occ = rng.uniform(0, 1, (M, 54)) < rng.uniform(0.2, 1.0, (M, 1))
    dem = np.where(occ, np.clip(rng.lognormal(1.5, 1.2, (M, 54)), 0.002, 100), 0).astype(np.float32)
    x, info = O.project_action(net, a, dem)
    slow = info["slow"]
    assert len(slow) > 20
    assert info["kkt"][slow].max() < 1e-10 and info["iters"][slow].max() < 40
    At, r, u = net.A_tilde, net.magnitudes / 32, O.projection_bounds(dem)
    assert (np.abs(x @ At.T) - r[None]).max() < 1e-12 and (x - u).max() <= 0 and x.min() >= 0
    fast = np.setdiff1d(np.arange(M), slow)
    np.testing.assert_array_equal(x[fast], np.clip(a[fast].astype(np.float64), 0, u[fast]))
    for m in slow[:8]:
        am, um = a[m].astype(float), u[m]
        cons = [{"type": "ineq", "fun": (lambda z, j=j: r[j] ** 2 - abs(At[j] @ z) ** 2),
                 "jac": (lambda z, j=j: -2 * ((At[j] @ z).real * At[j].real + (At[j] @ z).imag * At[j].imag))}
                for j in range(8)]
        res = minimize(lambda z: 0.5 * np.sum((z - am) ** 2), np.clip(am, 0, um) * 0.5, jac=lambda z: z - am,
                       bounds=list(zip(np.zeros(54), um)), constraints=cons, method="SLSQP",
                       options={"ftol": 1e-15, "maxiter": 500})
        assert np.abs(res.x - x[m]).max() < 1e-6


def test_trace_tables_roundtrip():
    days = synthetic_days(5, seed=2, mean_sessions=40)
    t = pack_sessions(days, 54)
    for d in range(5):
        assert t.sessions_of_day(d) == sorted(days[d], key=lambda s: (s[0], s[4]))
    with pytest.raises(ValueError):
        pack_sessions([[(10, 50, 40, 5.0, 3), (30, 60, 55, 5.0, 3)]], 54)   # overlap on one station
