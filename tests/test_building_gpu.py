"""GPU: BuildingEnv kernels vs (a) the reference's own trajectories (golden fixtures) and (b) the
NumPy oracle on seeded random batches.  Tolerances: observations (f32 state) 1e-5 relative as the
north star states (we assert 1e-6); cursor / counters / done bit-exact; reward 1e-6 relative."""
import numpy as np
import pytest
import torch

from helpers import building_fixtures, load_building_fixture, oracle_building_params, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", building_fixtures(), ids=lambda p: p.split("building_")[-1][:-4])
def test_kernel_matches_reference_trajectory(native_lib, path):
    from sustaingym_b200.envs.building.env import BuildingEnv
    g, params = load_building_fixture(path)
    env = BuildingEnv(params)
    obs0, _ = env.reset(seed=int(g["seed"]))
    assert env.epoch == int(g["epoch0"])
    np.testing.assert_array_equal(obs0, g["obs0"])
    exact = 0
    for t in range(len(g["actions"])):
        ob, r, term, trunc, info = env.step(g["actions"][t])
        exact += np.array_equal(ob, g["obs"][t])
        assert rel_err(ob, g["obs"][t]) < 1e-6, t
        assert abs(r - g["rewards"][t]) <= 1e-6 * abs(g["rewards"][t]), t
        assert term == trunc == bool(g["dones"][t])
    assert exact >= len(g["actions"]) - 3


@pytest.mark.parametrize("name,B", [("retail_s5", 1000), ("officesmall_occ", 777), ("officemedium_s9", 300),
                                    ("retail_discrete", 513)])
def test_kernel_matches_oracle_random_batch(native_lib, name, B):
    from oracle.building import BuildingOracle
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    path = [p for p in building_fixtures() if name in p][0]
    g, params = load_building_fixture(path)
    env = BatchedBuildingEnv(params, B)
    orc = BuildingOracle(oracle_building_params(g), B)
    rng = np.random.default_rng(7)
    T = len(g["out_temp"])
    epochs = rng.integers(0, T, B)
    epochs[:4] = [T - 1, T - 2, 0, T - 100]
    T0 = rng.uniform(15, 30, (B, env.n))
    obs, _ = env.reset(epochs=epochs, options={"T_initial": T0})
    o_obs = orc.reset(epochs, T0)
    np.testing.assert_array_equal(obs.cpu().numpy(), o_obs)
    n_exact = 0
    for t in range(60):
        if env.is_continuous_action:
            a = (rng.uniform(-1, 1, (B, env.n)) * env.ac_map).astype(np.float32)
        else:
            a = rng.integers(0, 200, (B, env.n)).astype(np.int64)
        ob, r, term, trunc, _ = env.step(torch.from_numpy(a).cuda())
        o_ob, o_r, o_d, _ = orc.step(a)
        ob = ob.cpu().numpy()
        n_exact += np.array_equal(ob, o_ob)
        assert rel_err(ob, o_ob) < 1e-6
        np.testing.assert_allclose(r.cpu().numpy(), o_r, rtol=1e-6)
        np.testing.assert_array_equal(term.cpu().numpy(), o_d)
        np.testing.assert_array_equal(env.epoch.cpu().numpy(), orc.epoch)
        np.testing.assert_array_equal(env.runs.cpu().numpy(), orc.runs)
        # keep the two in lock-step on the exact same f32 state
        orc.state = ob.copy()
    assert n_exact >= 50


def test_rollout_equals_repeated_step_and_host_path(native_lib):
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    g, params = load_building_fixture([p for p in building_fixtures() if "retail_s5" in p][0])
    B, T = 300, 40
    rng = np.random.default_rng(3)
    acts = rng.uniform(-1, 1, (T, B, 5)).astype(np.float32)
    epochs = rng.integers(0, 1900, B)
    e1, e2, e3 = (BatchedBuildingEnv(params, B) for _ in range(3))
    for e in (e1, e2, e3):
        e.reset(epochs=epochs)
    obs_r, rew_r, done_r = e1.rollout(torch.from_numpy(acts).cuda())
    for t in range(T):
        ob, r, d, _, _ = e2.step(torch.from_numpy(acts[t]).cuda())
        assert torch.equal(ob, obs_r[t]) and torch.equal(r, rew_r[t]) and torch.equal(d, done_r[t].bool())
        ob_h, r_h, d_h, _, _ = e3.step(acts[t])
        np.testing.assert_array_equal(ob_h, ob.cpu().numpy())
        np.testing.assert_array_equal(r_h, r.cpu().numpy())
    assert torch.equal(e1.state, e2.state) and torch.equal(e1.epoch, e2.epoch)


def test_autoreset_and_stats(native_lib):
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    g, params = load_building_fixture([p for p in building_fixtures() if "officesmall_s5" in p][0])
    params = dict(params, episode_len=10)
    B = 257
    env = BatchedBuildingEnv(params, B, auto_reset=True, collect_stats=True)
    env.reset(epochs=np.arange(B) % 1900)
    rng = np.random.default_rng(0)
    tot = 0.0
    for t in range(25):
        a = torch.from_numpy(rng.uniform(-1, 1, (B, 6)).astype(np.float32)).cuda()
        ob, r, d, _, _ = env.step(a)
        tot += float(r.sum())
        assert bool(d.all()) == (t % 10 == 9)
        if t % 10 == 9:
            assert torch.all(env.runs == 0) and torch.all(ob[:, :6] == 22.0)
    st = env.stats.cpu().numpy()
    assert st[0] == 25 * B and st[5] == 2 * B
    assert abs(st[1] - tot) < 1e-6 * abs(tot)


@pytest.mark.parametrize("name,B", [("outpatient_s3", 77), ("apartmenthighrise_s21", 64), ("hospital_s6", 33),
                                    ("officemedium_s9", 130)])
def test_tensor_core_kernel_matches_thread_per_env_kernel(native_lib, name, B):
    """n >= 16: DMMA cluster kernel (default) vs the thread-per-env kernel on the same inputs — fused rollout
    with auto-reset, per-step observations and statistics.  The two sum each dot product in a different
    order (f64), so the f32 states agree bit for bit except for rare last-bit roundings."""
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    g, params = load_building_fixture([p for p in building_fixtures() if name in p][0])
    params = dict(params, episode_len=20)
    S = 50
    rng = np.random.default_rng(3)
    epochs = rng.integers(0, len(g["out_temp"]), B)
    acts = torch.from_numpy((rng.uniform(-1, 1, (S, B, params["n"])) * np.asarray(params["ac_map"])).astype(np.float32)).cuda()
    outs = []
    for tc in (True, False):
        env = BatchedBuildingEnv(params, B, auto_reset=True, collect_stats=True, tensor_cores=tc)
        env.reset(epochs=epochs)
        obs, rew, done = env.rollout(acts, keep_obs=True)
        o1, r1, d1, _, _ = env.step(acts[0])                 # single-step launch after the rollout
        torch.cuda.synchronize()
        outs.append((obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), o1.cpu().numpy(), r1.cpu().numpy(),
                     env.epoch.cpu().numpy(), env.runs.cpu().numpy(), env.stats.cpu().numpy()))
    a, b = outs
    assert (a[0] != b[0]).mean() < 1e-4 and rel_err(a[0], b[0]) < 1e-6
    np.testing.assert_allclose(a[1], b[1], rtol=1e-9)
    np.testing.assert_array_equal(a[2], b[2])
    assert a[2].sum() == 2 * B                               # two episode ends per env in 50 steps
    assert rel_err(a[3], b[3]) < 1e-6
    np.testing.assert_allclose(a[4], b[4], rtol=1e-9)
    np.testing.assert_array_equal(a[5], b[5]); np.testing.assert_array_equal(a[6], b[6])
    np.testing.assert_allclose(a[7][:3], b[7][:3], rtol=1e-9)


def test_standalone_building_by_name_matches_survey_vectors(native_lib):
    """``BuildingEnv("RetailStandalone")`` — parameters from the packaged archive through our ParameterGenerator — against
    the vectors the survey recorded from the unmodified reference (Hot_Dry / Tucson, reset(seed=5)
    -> epoch 1439, actions = successive default_rng(0).uniform(-1, 1, n) as f32)."""
    from sustaingym_b200.envs.building.env import BuildingEnv
    env = BuildingEnv("RetailStandalone")
    obs, _ = env.reset(seed=5)
    assert env.epoch == 1439 and env.n == 5
    rng = np.random.default_rng(0)
    want = [([21.803447723388672, 21.775115966796875, 19.753095626831055, 19.665117263793945, 29.285329818725586, 11.0,
              20.899999618530273, 0.0, -0.07856952399015427], -8.008709221225727),
            ([22.030057907104492, 21.68138313293457, 20.758813858032227, 20.03554916381836, 34.25482940673828, 11.0,
              20.899999618530273, 0.0, -0.08379818499088287], -12.496131927911641)]
    for ob_w, r_w in want:
        ob, r, *_ = env.step(rng.uniform(-1, 1, 5).astype(np.float32))
        np.testing.assert_allclose(ob, np.array(ob_w, np.float32), rtol=1e-6, atol=1e-6)
        assert abs(r - r_w) <= 1e-6 * abs(r_w)
    env6 = BuildingEnv("OfficeSmall")
    obs, _ = env6.reset(seed=5)
    np.testing.assert_allclose(obs, np.array([22, 22, 22, 22, 22, 22, 11.0, 20.899999618530273, 0.0, -0.07856951653957367], np.float32), rtol=1e-6)


def test_mpc_agent_on_the_drop_in_env(native_lib):
    """mpc_controller.py's call sequence on the drop-in env: ``MPCAgent(env, beta, pnorm, planning_steps=..).predict()``
    reads env.A_d / B_d / out_temp / ... / state; the state it predicts is the one the kernel produces."""
    from sustaingym_b200.algorithms.building.mpc_controller import MPCAgent
    from sustaingym_b200.envs.building.env import BuildingEnv
    env = BuildingEnv("OfficeSmall")
    env.reset(seed=5, options={"T_initial": np.array([25.5, 24.0, 26.0, 27.0, 23.5, 25.0])})
    ag = MPCAgent(env, beta=float(env.parameters["reward_beta"]), pnorm=2, planning_steps=3, gap_tol=1e-6)
    total = 0.0
    for _ in range(12):
        a, pred = ag.predict()
        obs, r, *_ = env.step(a)
        np.testing.assert_allclose(obs[:env.n], pred, atol=2e-4)
        total += r
    idle = BuildingEnv("OfficeSmall")
    idle.reset(seed=5, options={"T_initial": np.array([25.5, 24.0, 26.0, 27.0, 23.5, 25.0])})
    assert total > sum(idle.step(np.zeros(idle.n, np.float32))[1] for _ in range(12))
