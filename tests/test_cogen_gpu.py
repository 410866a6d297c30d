"""GPU: CogenEnv kernels vs the oracle.  Plant-model tolerance: |d| <= 1e-5 *
max(|ref|, range_j) with range_j from model.json; rewards 1e-5 relative; obs/time/done exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.cogen import CogenOracle, PlantModel, flatten_obs

pytestmark = pytest.mark.gpu


def out_scale():
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "..", "sustaingym_b200", "data", "cogen_plant.npz"))
    span = z["out_max"] - z["out_min"]
    return np.where(np.isfinite(span) & (span > 0), span, 100.0)


@pytest.mark.parametrize("tensor_cores", [True, False])
def test_plant_operator_matches_onnx_golden(native_lib, tensor_cores):
    """tensor_cores=True: tcgen05 3xTF32 kernel (default); False: the fp32 FFMA kernel."""
    from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
    g = np.load(os.path.join(GOLDEN, "cogen_model_io.npz"))
    env = BatchedCogenEnv(1, tensor_cores=tensor_cores)
    Y = env.plant(torch.from_numpy(g["X"]).cuda()).cpu().numpy()
    scale = np.maximum(np.abs(g["Y32"]), out_scale()[None])
    err = np.abs(Y - g["Y32"]) / scale
    assert err.max() < 1e-5, err.max()
    # and against the f64 evaluation of the same graph, to show the f32 kernel is as good as f32 NumPy
    err64 = np.abs(Y - g["Y64"]) / np.maximum(np.abs(g["Y64"]), out_scale()[None])
    ref64 = np.abs(g["Y32"] - g["Y64"]) / np.maximum(np.abs(g["Y64"]), out_scale()[None])
    assert err64.max() < max(2 * ref64.max(), 1e-5)


@pytest.mark.parametrize("B,k", [(200, 12), (65, 3), (1, 0)])
def test_step_matches_oracle(native_lib, B, k):
    from sustaingym_b200.data.cogen_ambients import synthetic_ambients
    from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
    amb = synthetic_ambients(6, seed=3)
    env = BatchedCogenEnv(B, amb, forecast_horizon=k, keep_outputs=True)
    orc = CogenOracle(amb, B, forecast_horizon=k)
    rng = np.random.default_rng(1)
    days = rng.integers(0, 5, B)
    init = env.sample_actions(rng)
    obs, _ = env.reset(days=days, init_actions=init)
    o_obs = orc.reset(days, init)
    np.testing.assert_array_equal(obs.cpu().numpy(), flatten_obs(o_obs))
    scale = out_scale()[None]
    for t in range(96):
        a = env.sample_actions(rng)
        if t % 7 == 0:
            a[: B // 2, 2] = 0            # switch GT1 off for half the envs (gates)
        noise = rng.normal(size=(B, k, 7)) * 0.1 if k > 0 else None
        env.forecast_noise_std = 1.0
        ob, r, term, trunc, info = env.step(torch.from_numpy(a).cuda(), None if noise is None else torch.from_numpy(noise).cuda())
        o_ob, o_r, o_term, _, o_info = orc.step(a, noise)
        Y = env.Y.cpu().numpy()
        assert (np.abs(Y - o_info["Y"]) / np.maximum(np.abs(o_info["Y"]), scale)).max() < 1e-5
        # the reward is Lipschitz in the plant outputs: fuel terms enter with weight 1, the constraint hinges
        # max(0, Y_j - x) and the non-delivery terms with weight 1000 (the two penalties).  So a plant output
        # that is right to 1e-5 * range_j moves the reward by up to 1000 * |dY|_1 whatever the reward's size:
        # the bound below is rigorous given the measured output differences.
        bound = 1e-5 * np.abs(o_r) + 1001.0 * np.abs(Y.astype(np.float64) - o_info["Y"]).sum(axis=1) + 1e-3
        assert (np.abs(r.cpu().numpy() - o_r) <= bound).all()
        assert (np.abs(info["agent_rewards"].cpu().numpy() - o_info["agent_rewards"]) <= bound[:, None]).all()
        np.testing.assert_array_equal(ob.cpu().numpy(), flatten_obs(o_ob))          # obs: table look-ups, exact
        np.testing.assert_array_equal(term.cpu().numpy(), o_term)
    assert bool(term.all())


def test_single_and_multiagent_views(native_lib):
    from sustaingym_b200.data.cogen_ambients import synthetic_ambients
    from sustaingym_b200.envs.cogen.env import CogenEnv, MultiAgentCogenEnv
    amb = synthetic_ambients(8, seed=1)
    env = CogenEnv(ambients=amb)
    obs, info = env.reset(seed=3)
    assert env.observation_space.contains(obs) and env.current_day == 3
    tot = 0
    for t in range(96):
        obs, r, term, trunc, info = env.step(env.action_space.sample())
        assert isinstance(r, float) and env.observation_space.contains(obs) and trunc is False
        assert abs(sum(info["agent_rewards"].values()) - r) <= 1e-5 * abs(r)
    assert term
    ma = MultiAgentCogenEnv(ambients=amb, forecast_noise_std=0.1)
    obss, infos = ma.reset(seed=2)
    assert set(obss) == {"GT1", "GT2", "GT3", "ST"} and obss["GT1"].shape == (124,)
    for t in range(96):
        acts = {a: ma.action_space(a).sample() for a in ma.agents}
        obss, rews, terms, truncs, infos = ma.step(acts)
        assert all(obss[a].shape == (124,) for a in obss) and all(isinstance(v, float) for v in rews.values())
    assert ma.agents == [] and all(terms.values())
    assert ma.observation_space("GT1") is ma.observation_space("GT1")
    # seed reproducibility (parallel_seed_test): same seed -> same first observation
    o1, _ = ma.reset(seed=5); o2, _ = ma.reset(seed=5)
    np.testing.assert_array_equal(o1["ST"], o2["ST"])


def test_default_env_runs_on_the_reference_ambient_table(native_lib):
    """CogenEnv() without arguments uses the reference's own 253-day ambient table (packaged), like cogen/env.py:89."""
    from sustaingym_b200.envs.cogen.env import CogenEnv
    env = CogenEnv(renewables_magnitude=50.0)
    assert env._b.n_days == 253
    obs, _ = env.reset(seed=7)
    assert env.current_day == 7 and env.observation_space.contains(obs)
    for t in range(96):
        obs, r, term, trunc, info = env.step(env.action_space.sample())
        assert env.observation_space.contains(obs) and np.isfinite(r)
    assert term
