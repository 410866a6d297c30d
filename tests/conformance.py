"""API conformance checkers standing in for the three suites the reference's tests call
(``gymnasium.utils.env_checker.check_env``, ``pettingzoo.test.parallel_api_test`` /
``parallel_seed_test``, RLlib ``check_env``) — none of those libraries is installed here.
They assert the behaviours listed in SURVEY.md section 4."""
import numpy as np


def _same(a, b):
    if isinstance(a, dict):
        return set(a) == set(b) and all(_same(a[k], b[k]) for k in a)
    return np.array_equal(np.asarray(a), np.asarray(b))


def check_env(env, steps: int = 30):
    """gymnasium check_env essentials."""
    assert hasattr(env, "action_space") and hasattr(env, "observation_space")
    out = env.reset(seed=123)
    assert isinstance(out, tuple) and len(out) == 2 and isinstance(out[1], dict)
    obs1 = out[0]
    assert env.observation_space.contains(obs1), "reset obs not in observation_space"
    obs2, _ = env.reset(seed=123)
    assert _same(obs1, obs2), "reset(seed) is not deterministic"
    env.reset(seed=456)
    env.reset()                                          # seed=None accepted
    env.reset(seed=123, options={})
    env.action_space.seed(0)
    for _ in range(steps):
        a = env.action_space.sample()
        res = env.step(a)
        assert isinstance(res, tuple) and len(res) == 5
        obs, r, term, trunc, info = res
        assert env.observation_space.contains(obs), "step obs not in observation_space"
        assert isinstance(r, (float, np.floating)) and np.isfinite(r)
        assert isinstance(term, (bool, np.bool_)) and isinstance(trunc, (bool, np.bool_)) and isinstance(info, dict)
        if term or trunc:
            env.reset()
    # determinism of the dynamics under identical seeds and actions
    traj = []
    for _ in range(2):
        env.reset(seed=7)
        env.action_space.seed(7)
        t = []
        for _ in range(10):
            o, r, *_ = env.step(env.action_space.sample())
            t.append((o, r))
        traj.append(t)
    for (o1, r1), (o2, r2) in zip(*traj):
        assert _same(o1, o2) and r1 == r2


def parallel_api_test(env, num_cycles: int = 300, check_obs: bool = True):
    """pettingzoo.test.parallel_api_test essentials."""
    assert list(env.possible_agents) and isinstance(env.possible_agents, list)
    for a in env.possible_agents:
        assert env.observation_space(a) is env.observation_space(a)
        assert env.action_space(a) is env.action_space(a)
    obss, infos = env.reset()
    assert set(obss) == set(env.agents) == set(infos)
    cycles = 0
    while cycles < num_cycles:
        live = list(env.agents)
        acts = {a: env.action_space(a).sample() for a in live}
        obss, rews, terms, truncs, infos = env.step(acts)
        cycles += 1
        for d in (obss, rews, terms, truncs, infos):
            assert set(d) == set(live), "every live agent must appear in all five dicts"
        for a in live:
            if check_obs:
                assert env.observation_space(a).contains(np.asarray(obss[a], env.observation_space(a).dtype))
            assert isinstance(rews[a], (float, np.floating)) and isinstance(terms[a], (bool, np.bool_))
        finished = [a for a in live if terms[a] or truncs[a]]
        assert all(a not in env.agents for a in finished), "finished agents must leave env.agents"
        if not env.agents:
            obss, infos = env.reset()
            assert set(obss) == set(env.possible_agents)


def parallel_seed_test(factory, steps: int = 20):
    """Two envs, same seed, same seeded action streams -> identical observations and rewards."""
    e1, e2 = factory(), factory()
    o1, _ = e1.reset(seed=42)
    o2, _ = e2.reset(seed=42)
    assert _same(o1, o2)
    for a in e1.possible_agents:
        e1.action_space(a).seed(42)
        e2.action_space(a).seed(42)
    for _ in range(steps):
        if not e1.agents:
            break
        a1 = {a: e1.action_space(a).sample() for a in e1.agents}
        a2 = {a: e2.action_space(a).sample() for a in e2.agents}
        r1, r2 = e1.step(a1), e2.step(a2)
        assert _same(r1[0], r2[0]) and r1[1] == r2[1]
