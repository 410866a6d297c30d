"""GPU: EVChargingEnv kernels vs the CPU oracle (oracle/ev.py) on the same seeded inputs.

Bars: schedule (integer amps), est_departures, done, day cursor: bit-exact; demands / MOER /
reward: <= 1e-5 relative (asserted tighter).  With the projection on, a pilot may legitimately
differ when the exact projected current sits within 1e-6 A of a rounding tie (both solvers are
only accurate to ~1e-10); such envs are excluded from that step on and counted."""
import numpy as np
import pytest
import torch

from helpers import rel_err, tie_distance
from oracle import ev as O

pytestmark = pytest.mark.gpu


def make(num_days, mean_sessions, seed, B, project, **kw):
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    gen = SyntheticTraceGenerator("caltech", num_days=num_days, mean_sessions=mean_sessions, seed=seed)
    env = BatchedEVChargingEnv(gen, B, project_action_in_env=project, **kw)
    net = O.caltech_acn()
    sess = [[O.Session(*x) for x in d] for d in gen.days]
    orc = O.EVOracle(net, sess, gen.moer.astype(np.float64), B, moer_forecast_steps=env.moer_forecast_steps,
                     project_action_in_env=project)
    return gen, env, net, orc


def run_parity(env, orc, net, gen, B, steps, seed, action_fn, project):
    obs, _ = env.reset(seed=seed)
    days = gen.first_days(seed + np.arange(B), B)
    o_obs = orc.reset(days, gen.moer_days(days))
    np.testing.assert_array_equal(obs.cpu().numpy(), O.flatten_obs(o_obs))
    rng = np.random.default_rng(seed)
    clean = np.ones(B, bool)
    n_sched_exact = n_obs_exact = n_slow = 0
    for t in range(steps):
        a = action_fn(rng, t)
        ob, r, d, tr, info = env.step(torch.from_numpy(a).cuda())
        o_ob, o_r, o_d, _, o_info = orc.step(a if a.dtype == np.float32 else O.discrete_to_continuous(a, env.discrete_bins))
        n_slow += len(o_info["slow"])
        if project:
            clean &= tie_distance(net, o_info["x"]).min(axis=1) > 1e-6
        ob, r = ob.cpu().numpy(), r.cpu().numpy()
        f = O.flatten_obs(o_ob)
        sched = env.pilots.cpu().numpy().astype(np.float64)
        c = clean
        np.testing.assert_array_equal(sched[c], o_info["schedule"][c])
        n = net.n
        np.testing.assert_array_equal(ob[c, n:2 * n], f[c, n:2 * n])            # est_departures: integers
        np.testing.assert_array_equal(ob[c, 2 * n:], f[c, 2 * n:])              # MOER row + timestep
        assert rel_err(ob[c, :n], f[c, :n], floor=1e-3) < 1e-6                  # demands (kWh)
        np.testing.assert_allclose(r[c], o_r[c], rtol=1e-6, atol=1e-12)
        np.testing.assert_array_equal(d.cpu().numpy(), o_d)
        n_obs_exact += int(np.array_equal(ob[c], f[c]))
    bd = env.breakdown.cpu().numpy()
    np.testing.assert_allclose(bd[clean], orc.breakdown[clean], rtol=1e-6, atol=1e-12)
    return clean, n_obs_exact, n_slow


@pytest.mark.parametrize("B,mean", [(64, 25.0), (257, 70.0), (33, 120.0), (100, 60.0)])
def test_step_no_projection(native_lib, B, mean):
    gen, env, net, orc = make(7, mean, 11, B, False)
    clean, n_exact, _ = run_parity(env, orc, net, gen, B, 288, 3,
                                   lambda rng, t: rng.uniform(0, 1, (B, 54)).astype(np.float32), False)
    assert clean.all() and n_exact >= 250
    assert bool(env.done.all()) and int(env.t[0]) == 288


@pytest.mark.parametrize("B,mean", [(64, 25.0), (200, 80.0), (33, 120.0), (100, 100.0)])
def test_step_with_projection(native_lib, B, mean):
    gen, env, net, orc = make(5, mean, 5, B, True)
    clean, n_exact, n_slow = run_parity(env, orc, net, gen, B, 288, 9,
                                        lambda rng, t: rng.uniform(0, 1, (B, 54)).astype(np.float32), True)
    assert clean.mean() > 0.97
    if mean > 50:
        assert n_slow > 100            # the interior-point path was really exercised
    st = env.status.cpu().numpy()
    assert not (st & 2).any()          # no iteration-cap hits


def test_discrete_actions_config1(native_lib):
    """BASELINE config 1: discrete actions {0..4} -> a/4 (wrappers.py:43-45)."""
    B = 33
    gen, env, net, orc = make(3, 40.0, 2, B, True, discrete_bins=5)
    clean, _, _ = run_parity(env, orc, net, gen, B, 288, 1,
                             lambda rng, t: rng.integers(0, 5, (B, 54)).astype(np.int64), True)
    assert clean.mean() > 0.9


def test_projection_operator_matches_oracle(native_lib):
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    env = BatchedEVChargingEnv(SyntheticTraceGenerator(num_days=1), 1)
    net = O.caltech_acn()
    rng = np.random.default_rng(0)
    M = 2000
    a = rng.uniform(0, 1, (M, 54)).astype(np.float32)
    occ = rng.uniform(0, 1, (M, 54)) < rng.uniform(0.05, 1.0, (M, 1))
    dem = np.where(occ, np.clip(rng.lognormal(1.5, 1.2, (M, 54)), 0.002, 100), 0).astype(np.float32)
    a[0] = 0; dem[1] = 0; a[2] = 1; dem[2] = 100          # edge cases: zero action, empty garage, all-max
    x, st, it = env.project(torch.from_numpy(a).cuda(), torch.from_numpy(dem).cuda())
    xo, info = O.project_action(net, a, dem)
    x, st, it = x.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
    slow = np.zeros(M, bool); slow[info["slow"]] = True
    np.testing.assert_array_equal((st & 1).astype(bool), slow)
    assert not (st & 2).any() and it.max() < 40
    np.testing.assert_array_equal(x[~slow], xo[~slow])     # fast path: exact clip
    assert np.abs(x - xo).max() < 1e-9                     # solver path: both converge to the projection
    assert slow.sum() > 500


def test_autoreset_sequential_days_and_host_path(native_lib):
    B = 40
    gen, env, net, orc = make(4, 30.0, 8, B, False, auto_reset=True)
    _, env2, _, _ = make(4, 30.0, 8, B, False, auto_reset=True)
    env.reset(seed=0); env2.reset(seed=0)
    day0 = env.day.cpu().numpy().copy()
    rng = np.random.default_rng(0)
    for t in range(288 * 2 + 5):
        a = rng.uniform(0, 1, (B, 54)).astype(np.float32)
        ob, r, d, _, _ = env.step(torch.from_numpy(a).cuda())
        ob2, r2, d2, _, _ = env2.step(a)                                   # host-buffer path
        np.testing.assert_array_equal(ob.cpu().numpy(), ob2)
        np.testing.assert_array_equal(r.cpu().numpy(), r2)
        if t % 288 == 287:
            assert bool(d.all())
            assert torch.all(env.t == 0) and torch.all(ob[:, :108] == 0) and torch.all(ob[:, -1] == 0)
            assert env.final_obs[0, -1].item() == 1.0
        else:
            assert not bool(d.any())
    np.testing.assert_array_equal(env.day.cpu().numpy(), (day0 + 2) % 4)


def test_step_async_matches_step(native_lib):
    """step_async / step_wait (own stream, no sync inside the C call) == the synchronous host path."""
    B = 96
    _, env, _, _ = make(3, 60.0, 4, B, True)
    _, env2, _, _ = make(3, 60.0, 4, B, True)
    env.reset(seed=2); env2.reset(seed=2)
    rng = np.random.default_rng(1)
    for t in range(120):
        a = torch.from_numpy(rng.uniform(0, 1, (B, 54)).astype(np.float32)).pin_memory()
        env.step_async(a)
        ob, r, d, _, _ = env.step_wait()
        ob2, r2, d2, _, _ = env2.step(a.numpy())
        np.testing.assert_array_equal(ob.numpy(), ob2)
        np.testing.assert_array_equal(r.numpy(), r2)
        np.testing.assert_array_equal(d.numpy(), d2)


def test_export_state_matches_oracle(native_lib):
    """Dense view of the sparse in-place state: delivered kWh per station vs the oracle's sessions."""
    B = 50
    gen, env, net, orc = make(3, 60.0, 6, B, False)
    env.reset(seed=1)
    days = gen.first_days(1 + np.arange(B), B)
    orc.reset(days, gen.moer_days(days))
    rng = np.random.default_rng(3)
    for t in range(150):
        a = rng.uniform(0, 1, (B, 54)).astype(np.float32)
        env.step(torch.from_numpy(a).cuda()); orc.step(a)
    delivered, slot = env.export_state()
    delivered, slot = delivered.cpu().numpy(), slot.cpu().numpy()
    want = np.zeros((B, 54))
    bi, si = np.nonzero(orc.plugged)
    want[bi, orc._stn[orc.day][bi, si]] = orc.delivered[bi, si]
    occupied = np.zeros((B, 54), bool); occupied[bi, orc._stn[orc.day][bi, si]] = True
    np.testing.assert_array_equal(slot != 255, occupied)
    np.testing.assert_allclose(delivered, want, rtol=1e-9, atol=1e-12)


def test_single_env_gymnasium_view(native_lib):
    from sustaingym_b200.envs.evcharging.env import EVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    env = EVChargingEnv(SyntheticTraceGenerator(num_days=3, seed=4), project_action_in_env=True)
    obs, info = env.reset(seed=1)
    assert set(obs) == {"timestep", "est_departures", "demands", "prev_moer", "forecasted_moer"}
    assert env.observation_space.contains(obs)
    tot, done, steps = 0.0, False, 0
    while not done:
        obs, r, done, trunc, info = env.step(env.action_space.sample())
        assert isinstance(r, float) and trunc is False and env.observation_space.contains(obs)
        tot += r; steps += 1
    assert steps == 288
    bd = info["reward_breakdown"]
    assert abs(tot - (bd["profit"] - bd["carbon_cost"] - bd["excess_charge"])) < 1e-9
    obs_a, _ = env.reset(seed=1)
    obs_b, _ = env.reset(seed=1)
    for k in obs_a:
        np.testing.assert_array_equal(obs_a[k], obs_b[k])


@pytest.mark.parametrize("mean,cap,dense", [(30.0, None, 0), (110.0, None, 0), (110.0, 3, 0), (110.0, None, 97)])
def test_compact_host_path_is_bit_identical(native_lib, mean, cap, dense, monkeypatch):
    """Host-buffer path through 120-byte records + host expansion (sg_ev_step_host_compact_async, sg_ev_expand_obs)
    against the dense device->host copy: observations, rewards, done flags bit for bit, through an auto-reset;
    mean = 110 sessions/day puts more than 16 active sessions into many envs (overflow rows), cap = 3 forces the
    dense fallback of a step whose overflow area is too small."""
    B = 300
    monkeypatch.setenv("SG_HOST_DENSE_FRACTION", str(dense / B))      # dense = 97: the first 97 envs travel as dense rows
    gen, env_c, _, _ = make(4, mean, 8, B, True, auto_reset=True, host_obs="compact", host_threads=3)
    assert env_c.host_dense_envs == dense
    _, env_d, _, _ = make(4, mean, 8, B, True, auto_reset=True, host_obs="dense")
    env_c.reset(seed=0); env_d.reset(seed=0)
    if cap is not None:
        env_c._ensure_compact()["cap"] = cap
    rng = np.random.default_rng(0)
    n_over = 0
    for t in range(288 + 40):
        a = torch.from_numpy(rng.uniform(0, 1, (B, 54)).astype(np.float32)).pin_memory()
        if t % 3 == 0:
            env_c.step_async(a); oc, rc, dc, _, _ = env_c.step_wait()
            env_d.step_async(a); od, rd, dd, _, _ = env_d.step_wait()
            oc, rc, dc, od, rd, dd = (x.numpy() for x in (oc, rc, dc, od, rd, dd))
        else:
            oc, rc, dc, _, _ = env_c.step(a.numpy())
            od, rd, dd, _, _ = env_d.step(a.numpy())
        np.testing.assert_array_equal(oc, od, err_msg=f"step {t}")
        np.testing.assert_array_equal(rc, rd)
        np.testing.assert_array_equal(dc, dd)
        n_over += int(((od[:, :54] != 0).sum(1) > 16).sum())
        if t == 100:                      # a device-path step in between: the next host step must rewrite every row
            g = a.cuda()
            env_c.step(g); env_d.step(g)
    assert torch.equal(env_c.obs, env_d.obs)
    if mean > 100:
        assert n_over > 100               # the overflow rows were really used


def edge_case_days():
    """Hand-made days around the boundaries of the event model."""
    full = [(2, 280, 270, 60.0, i) for i in range(54)]                     # every station occupied all day long
    back_to_back = []
    for i in range(0, 54, 3):                                               # a departure and an arrival in the same period,
        t = 0                                                               # session after session on the same station
        while t + 9 < 287:
            back_to_back.append((t, t + 9, min(t + 9 + (i % 5) - 2, 287), 3.0 + i % 7, i))
            t += 9
    odd = [(0, 287, 287, 100.0, 0), (1, 287, 287, 150.0, 1),              # first / last periods, energy above the 100 kWh cap
           (5, 5, 6, 10.0, 2), (7, 8, 8, 0.0005, 3),                      # zero-length session; fully charged on arrival
           (10, 200, 12, 0.2, 4), (10, 200, 11, 90.0, 5),                 # estimated departure long passed; tiny / huge demand
           (285, 287, 287, 20.0, 6), (286, 287, 287, 20.0, 7), (287, 287, 287, 5.0, 8)]   # arrivals in the last periods
    pods = [(3, 250, 240, 40.0, i) for i in range(10, 26)] + [(100, 180, 170, 8.0, i) for i in range(26, 54)]   # both pods full
    return [[], full, back_to_back, odd, pods]


@pytest.mark.parametrize("project", [False, True])
def test_edge_case_days(native_lib, project):
    """Empty day, 54 of 54 stations occupied (one env fills a whole item group; every pod and winding constraint binds),
    sessions chained on one station, zero-length / instantly full / over-cap sessions, events in the first and last
    periods — kernel vs oracle over the whole day, several envs per day (different actions)."""
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import ArrayTraceGenerator, synthetic_moer
    days = edge_case_days()
    gen = ArrayTraceGenerator(days, synthetic_moer(len(days), 3))
    B = 40
    env = BatchedEVChargingEnv(gen, B, project_action_in_env=project)
    net = O.caltech_acn()
    # (the generator caps the requested energy at 100 kWh when it packs the sessions, event_generation.py:169-170; the
    # oracle takes sessions that are capped already)
    sess = [[O.Session(a, dp, e, min(kwh, 100.0), st) for a, dp, e, kwh, st in d] for d in gen.days]
    orc = O.EVOracle(net, sess, gen.moer.astype(np.float64), B, moer_forecast_steps=36, project_action_in_env=project)

    def actions(rng, t):
        a = rng.uniform(0, 1, (B, 54)).astype(np.float32)
        a[::4] = 1.0                                   # a quarter of the envs always ask for everything
        a[1::8] = 0.0                                  # ... and some for nothing
        return a
    clean, n_exact, n_slow = run_parity(env, orc, net, gen, B, 288, 0, actions, project)
    assert clean.mean() > 0.9
    if project:
        assert n_slow > 1000                           # the interior-point path ran (full garage, everything asked for)
