"""GPU: parity at the BASELINE batch sizes (BASELINE.json configs 2, 4, 5).

The small-batch parity tests never make a warp of ``ev_step_main_kernel`` loop over more than one tile of envs; the
benchmark batch (65 536 envs, 2.3 tiles per warp) does.  Here:

  * EV at B = 65 536 against the full-batch oracle over 40 steps of a day whose garage is busy from the first period
    (mid-day occupancy, projection on) — `env.py:229-291`;
  * EV at B = 65 536 with the grid capped (``SG_EV_MAX_CTAS``: >= 100 tiles per warp) against the uncapped launch,
    bit for bit, over more than one episode with auto-reset (ties the persistent loop to the oracle-checked path);
  * the small-batch oracle runs again under ``SG_EV_MAX_CTAS=1`` (every warp walks several tiles);
  * Building n = 5 at 32 768 envs (config 4's per-GPU shard) single step + fused rollout against the oracle;
  * Cogen at 4 096 (config 5's shard) and 32 768 envs against the oracle.
"""
import os

import numpy as np
import pytest
import torch

from helpers import building_fixtures, load_building_fixture, oracle_building_params, rel_err, tie_distance
from oracle import ev as O

pytestmark = pytest.mark.gpu


def busy_days(num_days: int, seed: int, n: int = 54, early: int = 12, later: float = 13.0) -> list:
    """Day traces whose garage holds ~`early` EVs from the first periods on (arrivals 0..3), plus Poisson(`later`)
    sessions over the rest of the day — the occupancy the benchmark sees around noon, reachable without a pre-roll."""
    rng = np.random.default_rng(seed)
    days = []
    for _ in range(num_days):
        free_at = np.full(n, -1)
        ss = []
        m = int(rng.poisson(later))
        arr = np.concatenate([rng.integers(0, 4, early), np.sort(rng.integers(4, 241, m))])
        for a in arr:
            avail = np.nonzero(free_at < a)[0]
            if len(avail) == 0:
                continue
            i = int(rng.choice(avail))
            dep = int(min(a + rng.integers(30, 200), 287))
            est = int(np.clip(dep + rng.integers(-12, 13), a + 1, 287))
            free_at[i] = dep
            ss.append((int(a), dep, est, float(np.clip(rng.lognormal(2.4, 0.8), 0.5, 100.0)), i))
        days.append(ss)
    return days


@pytest.fixture
def capped_grid():
    def setcap(v):
        if v is None:
            os.environ.pop("SG_EV_MAX_CTAS", None)
        else:
            os.environ["SG_EV_MAX_CTAS"] = str(v)
    yield setcap
    os.environ.pop("SG_EV_MAX_CTAS", None)


def test_ev_batch65536_matches_oracle(native_lib):
    """BASELINE config 2 at its stated batch: every output of 40 busy-garage steps against the oracle."""
    from sustaingym_b200.data.ev_traces import synthetic_moer
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import ArrayTraceGenerator
    B, D, steps = 65536, 64, 40
    days = busy_days(D, 5)
    gen = ArrayTraceGenerator(days, synthetic_moer(D, 6))
    env = BatchedEVChargingEnv(gen, B, project_action_in_env=True)
    net = O.caltech_acn()
    orc = O.EVOracle(net, [[O.Session(*x) for x in d] for d in days], gen.moer.astype(np.float64), B,
                     moer_forecast_steps=36, project_action_in_env=True)
    obs, _ = env.reset(seed=0)
    d0 = gen.first_days(np.arange(B), B)
    o_obs = orc.reset(d0, gen.moer_days(d0))
    np.testing.assert_array_equal(obs.cpu().numpy(), O.flatten_obs(o_obs))
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    clean = np.ones(B, bool)
    n_slow = 0
    occ_mean = 0.0
    for t in range(steps):
        # U(0, 0.75): the pods are over their limit in ~0.2 % of the envs per step (the exact projection runs, and
        # the SciPy side of the oracle stays within a minute)
        a = torch.rand(B, 54, generator=g, device="cuda", dtype=torch.float32) * 0.75
        ob, r, d, _, _ = env.step(a)
        o_ob, o_r, o_d, _, info = orc.step(a.cpu().numpy())
        n_slow += len(info["slow"])
        clean &= tie_distance(net, info["x"]).min(axis=1) > 1e-6
        ob, r = ob.cpu().numpy(), r.cpu().numpy()
        f = O.flatten_obs(o_ob)
        c = clean
        np.testing.assert_array_equal(env.pilots.cpu().numpy()[c], info["schedule"][c].astype(np.uint8))
        np.testing.assert_array_equal(ob[c, 54:], f[c, 54:])               # est_departures, MOER row, timestep: exact
        assert rel_err(ob[c, :54], f[c, :54], floor=1e-3) < 1e-6           # demands
        np.testing.assert_allclose(r[c], o_r[c], rtol=1e-6, atol=1e-12)
        np.testing.assert_array_equal(d.cpu().numpy(), o_d)
        occ_mean += (f[:, :54] > 0).sum(1).mean() / steps
    assert clean.mean() > 0.99
    assert occ_mean > 9.0 and n_slow > 200          # a busy garage, and the exact projection really ran
    np.testing.assert_allclose(env.breakdown.cpu().numpy()[clean], orc.breakdown[clean], rtol=1e-6, atol=1e-12)
    assert not (env.status.cpu().numpy() & 2).any()


def test_ev_capped_grid_is_bitwise_identical(native_lib, capped_grid):
    """B = 65 536, auto-reset, 300 steps: the launch with 16 CTAs (64 warps, 64 tiles each) writes exactly what the
    full grid writes (observations, rewards, pilots, state)."""
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    B = 65536
    gen = SyntheticTraceGenerator("caltech", num_days=64, mean_sessions=25.0, seed=3)
    envs = [BatchedEVChargingEnv(gen, B, project_action_in_env=True, auto_reset=True) for _ in range(2)]
    for e in envs:
        e.reset(seed=150)        # day index = seed + env id
    acts = [torch.rand(B, 54, device="cuda") for _ in range(4)]
    # fast-forward both to t = 100 (full grid), then compare capped vs uncapped from there through the day end
    for t in range(100):
        for e in envs:
            e.step(acts[t % 4])
    assert torch.equal(envs[0].obs, envs[1].obs)
    for t in range(100, 300):
        capped_grid(None)
        o0, r0, d0, _, _ = envs[0].step(acts[t % 4])
        capped_grid(16)
        o1, r1, d1, _, _ = envs[1].step(acts[t % 4])
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1), t
        assert torch.equal(envs[0].pilots, envs[1].pilots), t
        if t == 287:
            assert bool(d0.all()) and torch.equal(envs[0].final_obs, envs[1].final_obs)
    capped_grid(None)
    assert torch.equal(envs[0].header, envs[1].header)
    d_a, k_a = envs[0].export_state(); d_b, k_b = envs[1].export_state()
    assert torch.equal(d_a, d_b) and torch.equal(k_a, k_b)


@pytest.mark.parametrize("tiling", ["0,8", "0,6", "3,2", "5,10", "1,16", "0,1"])
def test_ev_short_tiles_are_bitwise_identical(native_lib, capped_grid, tiling, monkeypatch):
    """The launcher cuts the envs of a thin last round into short tiles (sg_ev_step: tiling).  Any tiling — forced here
    through SG_EV_TILING = "full 16-env tiles, envs per tile after them", with bulk and plain-store write-outs, odd
    sizes, a ragged end — writes exactly what the uniform tiling writes, through an end of day with auto-reset."""
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    B = 333
    gen = SyntheticTraceGenerator("caltech", num_days=8, mean_sessions=80.0, seed=4)
    envs = [BatchedEVChargingEnv(gen, B, project_action_in_env=True, auto_reset=True) for _ in range(2)]
    for e in envs:
        e.reset(seed=3)
        e.t[:] = 240                                  # busy afternoon -> end of day -> next morning
    capped_grid(2)
    acts = [torch.rand(B, 54, device="cuda") for _ in range(4)]
    for t in range(70):
        monkeypatch.delenv("SG_EV_TILING", raising=False)
        o0, r0, d0, _, _ = envs[0].step(acts[t % 4])
        monkeypatch.setenv("SG_EV_TILING", tiling)
        o1, r1, d1, _, _ = envs[1].step(acts[t % 4])
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1), t
        assert torch.equal(envs[0].pilots, envs[1].pilots) and torch.equal(envs[0].status, envs[1].status), t
    monkeypatch.delenv("SG_EV_TILING", raising=False)
    assert torch.equal(envs[0].header, envs[1].header) and torch.equal(envs[0].final_obs, envs[1].final_obs)
    d_a, k_a = envs[0].export_state(); d_b, k_b = envs[1].export_state()
    assert torch.equal(d_a, d_b) and torch.equal(k_a, k_b)


@pytest.mark.parametrize("B,mean,project", [(257, 70.0, True), (200, 80.0, True), (257, 70.0, False)])
def test_ev_oracle_parity_with_many_tiles_per_warp(native_lib, capped_grid, B, mean, project):
    """The 288-step oracle comparison of test_ev_gpu.py with ONE CTA: 4 warps share 13-17 tiles of 16 envs."""
    from test_ev_gpu import make, run_parity
    capped_grid(1)
    gen, env, net, orc = make(5, mean, 5, B, project)
    clean, n_exact, n_slow = run_parity(env, orc, net, gen, B, 288, 9,
                                        lambda rng, t: rng.uniform(0, 1, (B, 54)).astype(np.float32), project)
    assert clean.mean() > 0.97
    if project:
        assert n_slow > 100 and not (env.status.cpu().numpy() & 2).any()


def test_building_batch32768_matches_oracle(native_lib):
    """BASELINE config 4's per-GPU shard (262 144 / 8): single steps and the fused rollout against the oracle."""
    from oracle.building import BuildingOracle
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    g, params = load_building_fixture([p for p in building_fixtures() if "retail_s5" in p][0])
    B, T = 32768, 24
    rng = np.random.default_rng(11)
    epochs = rng.integers(0, len(g["out_temp"]), B)
    T0 = rng.uniform(15, 30, (B, 5))
    acts = (rng.uniform(-1, 1, (T, B, 5)) * np.asarray(params["ac_map"])).astype(np.float32)
    e_step, e_roll = BatchedBuildingEnv(params, B), BatchedBuildingEnv(params, B)
    orc = BuildingOracle(oracle_building_params(g), B)
    for e in (e_step, e_roll):
        e.reset(epochs=epochs, options={"T_initial": T0})
    orc.reset(epochs, T0)
    obs_r, rew_r, done_r = e_roll.rollout(torch.from_numpy(acts).cuda())
    n_exact = 0
    for t in range(T):
        ob, r, term, _, _ = e_step.step(torch.from_numpy(acts[t]).cuda())
        assert torch.equal(ob, obs_r[t]) and torch.equal(r, rew_r[t]) and torch.equal(term, done_r[t].bool())
        o_ob, o_r, o_d, _ = orc.step(acts[t])
        ob = ob.cpu().numpy()
        n_exact += np.array_equal(ob, o_ob)
        assert rel_err(ob, o_ob) < 1e-6
        np.testing.assert_allclose(r.cpu().numpy(), o_r, rtol=1e-6)
        np.testing.assert_array_equal(term.cpu().numpy(), o_d)
        np.testing.assert_array_equal(e_step.epoch.cpu().numpy(), orc.epoch)
        orc.state = ob.copy()
    assert n_exact >= T - 4


@pytest.mark.parametrize("B,steps", [(4096, 6), (32768, 2)])
def test_cogen_config5_batches_match_oracle(native_lib, B, steps):
    """BASELINE config 5: 32 768 envs (whole job) and 4 096 (one of 8 shards), k = 12, 4 agents."""
    from oracle.cogen import CogenOracle, flatten_obs
    from sustaingym_b200.data.cogen_ambients import synthetic_ambients
    from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
    from test_cogen_gpu import out_scale
    amb = synthetic_ambients(16, seed=5)
    env = BatchedCogenEnv(B, amb, forecast_horizon=12, keep_outputs=True)
    orc = CogenOracle(amb, B, forecast_horizon=12)
    rng = np.random.default_rng(2)
    days = rng.integers(0, 15, B)
    init = env.sample_actions(rng)
    obs, _ = env.reset(days=days, init_actions=init)
    np.testing.assert_array_equal(obs.cpu().numpy(), flatten_obs(orc.reset(days, init)))
    scale = out_scale()[None]
    for t in range(steps):
        a = env.sample_actions(rng)
        a[::3, 2] = 0
        ob, r, term, _, info = env.step(torch.from_numpy(a).cuda())
        o_ob, o_r, o_term, _, o_info = orc.step(a, None)
        Y = env.Y.cpu().numpy()
        assert (np.abs(Y - o_info["Y"]) / np.maximum(np.abs(o_info["Y"]), scale)).max() < 1e-5
        bound = 1e-5 * np.abs(o_r) + 1001.0 * np.abs(Y.astype(np.float64) - o_info["Y"]).sum(axis=1) + 1e-3
        assert (np.abs(r.cpu().numpy() - o_r) <= bound).all()
        assert (np.abs(info["agent_rewards"].cpu().numpy() - o_info["agent_rewards"]) <= bound[:, None]).all()
        np.testing.assert_array_equal(ob.cpu().numpy(), flatten_obs(o_ob))
        np.testing.assert_array_equal(term.cpu().numpy(), o_term)
