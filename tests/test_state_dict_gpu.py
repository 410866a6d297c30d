"""GPU: checkpoint / resume of the batched envs (``envs/_state.py``): a state dict taken mid-episode
and loaded back — into the same env or into a freshly built one — reproduces the continuation bit for bit."""
import numpy as np
import pytest
import torch

from helpers import building_fixtures, load_building_fixture

pytestmark = pytest.mark.gpu


def continuation(env, actions, steps):
    out = []
    for t in range(steps):
        o, r, d, *_ = env.step(actions[t % len(actions)])
        out.append((o.clone(), r.clone(), d.clone()))
    return out


def assert_same(a, b):
    assert len(a) == len(b)
    for t, ((o0, r0, d0), (o1, r1, d1)) in enumerate(zip(a, b)):
        assert torch.equal(o0, o1) and torch.equal(r0, r1) and torch.equal(d0, d1), t


def roundtrip(make, actions, warm, steps, reset):
    env = make()
    reset(env)
    continuation(env, actions, warm)
    sd = env.state_dict()
    assert all(v.device.type == "cpu" for v in sd.values() if isinstance(v, torch.Tensor))
    first = continuation(env, actions, steps)
    env.load_state_dict(sd)                                   # rewind the same env
    assert_same(first, continuation(env, actions, steps))
    fresh = make()                                            # and a new env built with the same arguments
    reset(fresh)
    fresh.load_state_dict(sd)
    assert_same(first, continuation(fresh, actions, steps))
    return env, sd


def test_ev_state_dict(native_lib):
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    B = 300
    gen = SyntheticTraceGenerator("caltech", num_days=6, mean_sessions=80.0, seed=2)
    acts = [torch.rand(B, 54, device="cuda") for _ in range(5)]
    make = lambda: BatchedEVChargingEnv(gen, B, project_action_in_env=True, auto_reset=True, collect_stats=True)   # noqa: E731
    env, sd = roundtrip(make, acts, warm=200, steps=120, reset=lambda e: e.reset(seed=1))    # crosses the end of the day
    assert set(sd) >= {"slots", "header", "obs", "pilots", "stats"}
    other = BatchedEVChargingEnv(gen, B + 1, project_action_in_env=True, auto_reset=True, collect_stats=True)
    with pytest.raises(ValueError):
        other.load_state_dict(sd)


def test_building_state_dict(native_lib):
    from sustaingym_b200.envs.building.env import BatchedBuildingEnv
    _, params = load_building_fixture([p for p in building_fixtures() if "retail_s5" in p][0])
    B = 700
    acts = [torch.rand(B, params["n"], device="cuda") * 2 - 1 for _ in range(3)]
    make = lambda: BatchedBuildingEnv(dict(params, episode_len=40), B, auto_reset=True, collect_stats=True)       # noqa: E731
    roundtrip(make, acts, warm=25, steps=30, reset=lambda e: e.reset(epochs=np.arange(B) % 500))


def test_cogen_state_dict_restores_the_noise_stream(native_lib):
    from sustaingym_b200.envs.cogen.env import BatchedCogenEnv
    B = 128
    make = lambda: BatchedCogenEnv(B, forecast_horizon=3, forecast_noise_std=0.1)                                  # noqa: E731
    rng = np.random.default_rng(0)
    probe = make()
    acts = [torch.from_numpy(probe.sample_actions(rng)).cuda() for _ in range(3)]
    init = probe.sample_actions(rng)
    roundtrip(make, acts, warm=5, steps=6, reset=lambda e: e.reset(days=np.arange(B) % 50, init_actions=init))


def test_market_state_dict(native_lib):
    from sustaingym_b200.envs.electricitymarket.env import BatchedElectricityMarketEnv
    B = 500
    acts = [torch.rand(B, 2, device="cuda") * 100 for _ in range(3)]
    roundtrip(lambda: BatchedElectricityMarketEnv(B), acts, warm=50, steps=40, reset=lambda e: e.reset(seed=0))
