"""GPU: the reference's API-conformance test strategy (tests/test_evcharging.py:15-47,
tests/test_building.py:15-48, tests/test_cogen.py:14-43 of the reference) applied to the drop-in
single-env and multi-agent views, plus the vector-env adapters."""
import numpy as np
import pytest
import torch

from conformance import check_env, parallel_api_test, parallel_seed_test
from helpers import building_fixtures, load_building_fixture

pytestmark = pytest.mark.gpu


def ev_gen(seed=0, days=4, mean=40.0):
    from sustaingym_b200.envs.evcharging.event_generation import SyntheticTraceGenerator
    return SyntheticTraceGenerator("caltech", num_days=days, mean_sessions=mean, seed=seed)


def building_params():
    g, params = load_building_fixture([p for p in building_fixtures() if "officesmall_s5" in p][0])
    return params


def test_ev_single_agent_check_env(native_lib):
    from sustaingym_b200.envs.evcharging.env import EVChargingEnv
    check_env(EVChargingEnv(ev_gen(), project_action_in_env=True))
    check_env(EVChargingEnv(ev_gen(), project_action_in_env=False, discrete_bins=5))      # BASELINE config 1


def test_ev_multi_agent_api(native_lib):
    from sustaingym_b200.envs.evcharging.multiagent_env import MultiAgentEVChargingEnv
    parallel_api_test(MultiAgentEVChargingEnv(ev_gen(), project_action_in_env=False), num_cycles=600)
    parallel_api_test(MultiAgentEVChargingEnv(ev_gen(), periods_delay=3, discrete_bins=5), num_cycles=300)
    parallel_seed_test(lambda: MultiAgentEVChargingEnv(ev_gen(), periods_delay=2))


def test_ev_periods_delay_documented_semantics(native_lib):
    """delay_semantics="documented": the other stations' demands / est_departures are `periods_delay` steps old, own
    entries, MOER fields and timestep current (what multiagent_env.py:102-148 describes)."""
    from sustaingym_b200.envs.evcharging.multiagent_env import BatchedMultiAgentEVChargingEnv
    B, delay = 16, 3
    env = BatchedMultiAgentEVChargingEnv(ev_gen(mean=70.0), B, periods_delay=delay, project_action_in_env=False,
                                         delay_semantics="documented")
    env.reset(seed=0)
    hist = [env.env.obs.clone()]
    rng = np.random.default_rng(0)
    for t in range(60):
        env.step(torch.from_numpy(rng.uniform(0, 1, (B, 54)).astype(np.float32)).cuda())
        hist.append(env.env.obs.clone())
        per = env.agent_observations()
        cur, past = hist[-1], hist[max(0, len(hist) - 1 - delay)]
        exp = cur[:, None, :].repeat(1, 54, 1)
        exp[:, :, :108] = past[:, None, :108]
        idx = torch.arange(54)
        exp[:, idx, idx] = cur[:, idx]
        exp[:, idx, 54 + idx] = cur[:, 54 + idx]
        assert torch.equal(per, exp), t


def test_ev_periods_delay_reference_behaviour(native_lib):
    """Default: what the reference RETURNS.  Its single-agent env hands out one dict whose arrays it mutates in place
    (env.py:157-163, 383-394) and the multi-agent wrapper keeps that very dict in its deque (multiagent_env.py:119-139),
    so the 'delayed' arrays alias the current ones.  The wrapper logic is transcribed below around such a dict and
    driven with our env's observations: every agent gets the current aggregate observation."""
    from collections import deque
    from sustaingym_b200.envs.evcharging.multiagent_env import MultiAgentEVChargingEnv
    delay = 3
    ma = MultiAgentEVChargingEnv(ev_gen(mean=70.0), periods_delay=delay, project_action_in_env=False)
    keys = ("demands", "est_departures", "forecasted_moer", "prev_moer", "timestep")          # sorted: flatten order
    sl = ma._b.env.obs_slices
    live = {k: np.zeros(sl[k].stop - sl[k].start, np.float32) for k in keys}                     # the env's self._obs

    def env_get_observation(flat):                   # env.py:383-394: fill the SAME arrays, return the SAME dict
        for k in keys:
            live[k][:] = flat[sl[k]]
        return live

    past = deque(maxlen=delay)

    def create_dict_from_obs_agg(obs_agg, init):     # multiagent_env.py:102-148, verbatim logic
        flatten = lambda o: np.concatenate([o[k] for k in keys])                                 # noqa: E731
        if init:
            past.clear()
            for _ in range(delay):
                past.append(obs_agg)
            return {a: flatten(obs_agg) for a in ma.agents}
        first = past.popleft()
        past.append(obs_agg)
        td = {a: obs_agg.copy() for a in ma.agents}
        for i, a in enumerate(ma.agents):
            for var in ("est_departures", "demands"):
                td[a][var] = first[var]
                td[a][var][i] = obs_agg[var][i]
        return {a: flatten(td[a]) for a in ma.agents}

    obss, _ = ma.reset(seed=1)
    ref = create_dict_from_obs_agg(env_get_observation(ma._b.env.obs.cpu().numpy()[0]), init=True)
    rng = np.random.default_rng(2)
    n_busy = 0
    for t in range(80):
        for a in ma.agents:
            np.testing.assert_array_equal(obss[a], ref[a])
        obss, *_ = ma.step({a: rng.uniform(0, 1, 1).astype(np.float32) for a in ma.agents})
        ref = create_dict_from_obs_agg(env_get_observation(ma._b.env.obs.cpu().numpy()[0]), init=False)
        n_busy += int((obss[ma.agents[0]][:54] > 0).sum() > 0)
    assert n_busy > 20                                # the garage was not empty: the comparison is not vacuous


def test_building_api(native_lib):
    from sustaingym_b200.envs.building.env import BuildingEnv
    from sustaingym_b200.envs.building.multiagent_env import MultiAgentBuildingEnv
    p = building_params()
    env = BuildingEnv(p)
    # the reference's own observations fall outside its declared space:
    # check the API contract with widened bounds, values are pinned by the golden-trajectory tests
    env.observation_space.low[:] = -1e6
    env.observation_space.high[:] = 1e6
    check_env(env)
    ma = MultiAgentBuildingEnv(p)
    for sp in ma.observation_spaces.values():
        sp.low[:] = -1e6; sp.high[:] = 1e6
    parallel_api_test(ma, num_cycles=320)
    parallel_seed_test(lambda: MultiAgentBuildingEnv(p))


def test_cogen_api(native_lib):
    from sustaingym_b200.data.cogen_ambients import synthetic_ambients
    from sustaingym_b200.envs.cogen.env import CogenEnv, MultiAgentCogenEnv
    amb = synthetic_ambients(8, seed=0)
    check_env(CogenEnv(ambients=amb))
    parallel_api_test(MultiAgentCogenEnv(ambients=amb), num_cycles=200, check_obs=False)
    parallel_seed_test(lambda: MultiAgentCogenEnv(ambients=amb, forecast_noise_std=0.0))


def test_vector_adapters_ev(native_lib):
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.vector import GymnasiumVectorEnv, RLlibVectorEnv, SB3VecEnv
    N = 8
    mk = lambda: BatchedEVChargingEnv(ev_gen(), N, project_action_in_env=True)   # noqa: E731
    v = SB3VecEnv(mk())
    obs = v.reset()
    assert set(obs) == {"timestep", "est_departures", "demands", "prev_moer", "forecasted_moer"}
    assert obs["demands"].shape == (N, 54) and obs["timestep"].shape == (N, 1)
    rng = np.random.default_rng(0)
    n_done = 0
    for t in range(288 + 5):
        obs, rews, dones, infos = v.step(rng.uniform(0, 1, (N, 54)).astype(np.float32))
        assert rews.dtype == np.float32 and dones.shape == (N,) and len(infos) == N
        if dones.any():
            n_done += 1
            assert t == 287 and dones.all()
            assert infos[0]["terminal_observation"]["timestep"][0] == 1.0 and infos[0]["TimeLimit.truncated"] is False
            assert np.all(obs["timestep"] == 0)                       # auto-reset observation
    assert n_done == 1
    g = GymnasiumVectorEnv(mk())
    obs, infos = g.reset(seed=3)
    for t in range(288):
        obs, r, term, trunc, infos = g.step(rng.uniform(0, 1, (N, 54)).astype(np.float32))
    assert term.all() and "final_observation" in infos and infos["_final_observation"].all()
    r_ = RLlibVectorEnv(mk())
    obs_list, _ = r_.vector_reset()
    assert len(obs_list) == N and obs_list[0]["demands"].shape == (54,)
    o, rw, te, tr, inf = r_.vector_step(rng.uniform(0, 1, (N, 54)).astype(np.float32))
    assert len(o) == len(rw) == len(te) == N and isinstance(rw[0], float)
    o1, _ = r_.reset_at(2)
    assert o1["timestep"][0] == 0.0 and r_.env.t.cpu().numpy().tolist().count(0) == 1


def test_discrete_action_wrapper(native_lib):
    from sustaingym_b200.envs.evcharging.env import EVChargingEnv
    from sustaingym_b200.envs.wrappers import DiscreteActionWrapper
    e1 = DiscreteActionWrapper(EVChargingEnv(ev_gen(), project_action_in_env=True), bins=5)
    e2 = EVChargingEnv(ev_gen(), project_action_in_env=True, discrete_bins=5)       # fused in the kernel
    assert e1.action_space.nvec.tolist() == [5] * 54
    e1.reset(seed=1); e2.reset(seed=1)
    rng = np.random.default_rng(0)
    for _ in range(100):
        a = rng.integers(0, 5, 54)
        o1, r1, *_ = e1.step(a)
        o2, r2, *_ = e2.step(a)
        assert r1 == r2 and all(np.array_equal(o1[k], o2[k]) for k in o1)


def test_sb3_adapter_pipelined_path_matches_generic_path(native_lib):
    """SB3VecEnv over an auto-resetting engine (step_async enqueues on the env's stream, the kernel resets, compact
    records cross PCIe) returns what the generic adapter path (synchronous step + host-driven reset) returns; the
    infos list is reused and only finished envs get a dict."""
    import torch
    from sustaingym_b200.envs.evcharging.env import BatchedEVChargingEnv
    from sustaingym_b200.envs.vector import SB3VecEnv
    N = 40
    fast = SB3VecEnv(BatchedEVChargingEnv(ev_gen(), N, project_action_in_env=True, auto_reset=True), copy=False)
    slow = SB3VecEnv(BatchedEVChargingEnv(ev_gen(), N, project_action_in_env=True, host_obs="dense"))
    assert fast._fast and not slow._fast
    o_f, o_s = fast.reset(), slow.reset()
    rng = np.random.default_rng(1)
    pinned = torch.zeros(N, 54).pin_memory()
    infos_id = None
    for t in range(288 + 20):
        a = rng.uniform(0, 1, (N, 54)).astype(np.float32)
        if t % 2:
            pinned.numpy()[...] = a
            fast.step_async(pinned)                       # pinned torch tensor: no staging copy
        else:
            fast.step_async(a)
        o_f, r_f, d_f, i_f = fast.step_wait()
        infos_id = infos_id or id(i_f)
        assert id(i_f) == infos_id                        # one list, reused
        if t > 287:       # second episode: the kernel moved every env to its next day (the generic path re-seeds instead)
            assert not d_f.any() and all(len(x) == 0 for x in i_f) and np.all(o_f["timestep"] == np.float32((t - 287) / 288))
            continue
        slow.step_async(a)
        o_s, r_s, d_s, i_s = slow.step_wait()
        np.testing.assert_array_equal(r_f, r_s)
        np.testing.assert_array_equal(d_f, d_s)
        if t == 287:
            assert d_f.all() and np.all(o_f["timestep"] == 0) and np.all(o_f["demands"] == 0)
            for k in o_s:
                np.testing.assert_array_equal(i_f[3]["terminal_observation"][k], i_s[3]["terminal_observation"][k])
        else:
            for k in o_s:
                np.testing.assert_array_equal(o_f[k], o_s[k], err_msg=f"{k} at step {t}")
            assert not d_f.any() and all(len(x) == 0 for x in i_f)
