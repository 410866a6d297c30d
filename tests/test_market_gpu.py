"""GPU: ElectricityMarketEnv kernel vs oracle/market.py (spec-defined env, SPEC.md)."""
import numpy as np
import pytest
import torch

from oracle.market import MarketOracle, MarketParams

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("intermediate,k", [(True, 1), (False, 2)])
def test_step_matches_oracle(native_lib, intermediate, k):
    from sustaingym_b200.envs.electricitymarket.env import BatchedElectricityMarketEnv, default_fleet, synthetic_load_moer
    B = 300
    cost, cap = default_fleet()
    load, moer = synthetic_load_moer(5, k, seed=2)
    env = BatchedElectricityMarketEnv(B, load, moer, cost, cap, k=k, use_intermediate_rewards=intermediate)
    orc = MarketOracle(MarketParams(np.sort(cost), cap[np.argsort(cost, kind="stable")], k=k,
                                    use_intermediate_rewards=intermediate), load, moer, B)
    rng = np.random.default_rng(0)
    days = rng.integers(0, 5, B)
    obs, _ = env.reset(days=days)
    np.testing.assert_array_equal(obs.cpu().numpy(), orc.reset(days))
    for t in range(288):
        a = rng.uniform(0, 120, (B, 2 * k)).astype(np.float32)
        a[rng.uniform(size=B) < 0.2, k] = 0.0                 # some envs dump energy at any price
        a[rng.uniform(size=B) < 0.2, 0] = 500.0               # some charge at any price
        ob, r, term, trunc, info = env.step(torch.from_numpy(a).cuda())
        o_ob, o_r, o_term, _, o_info = orc.step(a)
        np.testing.assert_allclose(ob.cpu().numpy(), o_ob, rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(r.cpu().numpy(), o_r, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(info["price"].cpu().numpy(), o_info["price"], rtol=1e-12)
        np.testing.assert_allclose(info["energy"].cpu().numpy(), o_info["energy"], rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(term.cpu().numpy(), o_term)
    assert bool(term.all())
    e = info["energy"].cpu().numpy()
    assert e.min() >= 0 and e.max() <= 80 and e.std() > 1      # the battery really cycled


def test_single_env_and_discrete_wrapper(native_lib):
    from conformance import check_env
    from sustaingym_b200.envs.electricitymarket.env import DiscreteMarketWrapper, ElectricityMarketEnv
    env = ElectricityMarketEnv(month="2021-05", use_intermediate_rewards=False)
    check_env(env)
    w = DiscreteMarketWrapper(ElectricityMarketEnv())
    obs, _ = w.reset(seed=1)
    e0 = obs[1]
    for _ in range(10):
        obs, r, *_ = w.step(0)                                 # charge
    assert obs[1] > e0
    for _ in range(30):
        obs, r, *_ = w.step(2)                                 # discharge
    assert obs[1] < e0
