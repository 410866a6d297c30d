"""GPU: the GMM day-trace sampler kernel (sg_ev_sample_gmm_days) vs its oracle — bit-exact tables — and the
EV env stepping on a pool sampled on the device."""
import numpy as np
import pytest
import torch

from oracle import ev as O
from oracle import ev_gmm as G

pytestmark = pytest.mark.gpu

REC = np.dtype([("arr", "<i2"), ("dep", "<i2"), ("est", "<i2"), ("res", "<i2"), ("energy", "<f8")])
KWH_PER_STEP = 32 * (5 / 60 * 0.208)


def device_days(gmm, n_stations, D, first_day, seed, cap, K, smax):
    from sustaingym_b200.envs.evcharging.event_generation import GmmDevice
    dev = GmmDevice(gmm, n_stations, "cuda")
    pool = dev.alloc_pool(D, K, smax)
    dev.sample(pool, first_day, seed, cap, KWH_PER_STEP)
    torch.cuda.synchronize()
    rec = pool["sessions"].cpu().numpy().view(REC).reshape(D, K, n_stations)
    return rec, pool["arrivals"].cpu().numpy().view(np.uint32), pool["dropped"].cpu().numpy(), pool["max_kwh"].cpu().numpy()


@pytest.mark.parametrize("site,period,n_st,D,first,seed,cap,K,smax", [
    ("caltech", "Summer 2019", 54, 48, 0, 1, 100.0, 8, 128),
    ("caltech", "Summer 2021", 54, 33, 1000, 2**40 + 5, 20.0, 8, 128),
    ("caltech", "Fall 2019", 54, 40, 7, 3, 100.0, 1, 16),          # tiny tables: sessions are dropped
    ("jpl", "Summer 2019", 52, 24, 0, 9, 100.0, 8, 128),           # 52 stations, ~47 sessions/day
])
def test_kernel_matches_oracle_bit_exact(native_lib, site, period, n_st, D, first, seed, cap, K, smax):
    from sustaingym_b200.envs.evcharging.event_generation import load_gmm
    gmm = load_gmm(site, period)
    rec, arrivals, dropped, max_kwh = device_days(gmm, n_st, D, first, seed, cap, K, smax)
    o_rec, o_arr, o_drop, o_kwh = G.sample_days(gmm, seed, D, n_st, cap, K, smax, KWH_PER_STEP, first_day=first)
    for f in ("arr", "dep", "est"):
        np.testing.assert_array_equal(rec[f], o_rec[f])
    np.testing.assert_array_equal(rec["energy"].view(np.uint64), o_rec["energy"].view(np.uint64))      # f64 bit patterns
    np.testing.assert_array_equal(arrivals, o_arr)
    np.testing.assert_array_equal(dropped, o_drop)
    np.testing.assert_array_equal(max_kwh.view(np.uint64), o_kwh.view(np.uint64))
    if K == 1:
        assert dropped.sum() > 0
    assert (rec["arr"] != G.EMPTY).sum() > 5 * D


def test_env_steps_on_device_sampled_pool(native_lib):
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.evcharging.event_generation import GMMsTraceGenerator
    B = 96
    gen = GMMsTraceGenerator("caltech", "Summer 2019", seed=4, num_days=16, sequential=True, moer="synthetic")
    env = BatchedEVChargingEnv(gen, B, project_action_in_env=False)
    net = O.caltech_acn()

    def episode(seed):
        t = gen.tables
        sess = [[O.Session(*x) for x in t.sessions_of_day(d)] for d in range(t.num_days)]
        orc = O.EVOracle(net, sess, gen.moer.astype(np.float64), B, moer_forecast_steps=36, project_action_in_env=False)
        obs, _ = env.reset(seed=seed)
        days = gen.first_days(seed + np.arange(B), B)
        o_obs = orc.reset(days, gen.moer_days(days))
        np.testing.assert_array_equal(obs.cpu().numpy(), O.flatten_obs(o_obs))
        rng = np.random.default_rng(seed)
        for t_ in range(288):
            a = rng.uniform(0, 1, (B, 54)).astype(np.float32)
            ob, r, d, tr, info = env.step(torch.from_numpy(a).cuda())
            o_ob, o_r, o_d, _, o_info = orc.step(a)
            np.testing.assert_array_equal(env.pilots.cpu().numpy().astype(np.float64), o_info["schedule"])
            f = O.flatten_obs(o_ob)
            np.testing.assert_array_equal(ob.cpu().numpy()[:, 54:], f[:, 54:])
            np.testing.assert_allclose(ob.cpu().numpy()[:, :54], f[:, :54], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(r.cpu().numpy(), o_r, rtol=1e-6, atol=1e-12)
        assert bool(env.done.all())
        return t.arrival.copy()

    a0 = episode(2)
    mp = env._max_profit_per_day
    np.testing.assert_allclose(mp, gen.tables.max_profit(), rtol=1e-12)
    ptr = gen.device_pool["sessions"].data_ptr()
    gen.resample()                               # next 16 days of the stream, same buffers
    assert gen.device_pool["sessions"].data_ptr() == ptr
    a1 = episode(5)
    assert not np.array_equal(a0, a1)
    o_rec, *_ = G.sample_days(gen.gmm, 4, 16, 54, 100.0, 8, 128, KWH_PER_STEP, first_day=16)
    np.testing.assert_array_equal(a1, o_rec["arr"])


def test_argument_checks(native_lib):
    from sustaingym_b200.envs.evcharging.event_generation import GmmDevice, load_gmm
    dev = GmmDevice(load_gmm("caltech", "Summer 2019"), 54, "cuda")
    pool = dev.alloc_pool(4, 8, 128)
    pool["max_sessions"] = 0
    with pytest.raises(RuntimeError, match="max_sessions"):
        dev.sample(pool, 0, 1, 100.0, KWH_PER_STEP)
