"""Vector-environment views of the batched engines for RL libraries.

The reference is consumed through StableBaselines3 ``SubprocVecEnv`` (4 processes,
``examples/evcharging/train_stable_baselines.py:264-275``), RLlib auto-vectorisation
(``examples/evcharging/train_rllib.py:155-160``) and plain Gymnasium loops; here one in-process
kernel launch replaces the process pool.  None of those libraries is installed in this image, so the
adapters are duck-typed against their documented protocols:

* ``SB3VecEnv``        ``num_envs, observation_space, action_space, reset() -> obs,
                         step_async(actions), step_wait() -> (obs, rews f32[N], dones bool[N], infos)``,
                         auto-reset with ``infos[i]['terminal_observation']`` / ``'TimeLimit.truncated'``
* ``GymnasiumVectorEnv`` ``single_observation_space, single_action_space, reset(seed) -> (obs, infos),
                         step(actions) -> (obs, rew, terminated, truncated, infos)`` with
                         ``final_observation`` / ``final_info`` on auto-reset (gymnasium 0.28 semantics)
* ``RLlibVectorEnv``   ``vector_reset(), reset_at(i), vector_step(actions) -> 5 lists``

They work with any engine exposing ``num_envs, single_observation_space, single_action_space,
reset(seed=, mask=), step(np.ndarray)`` and a flat f32 observation (EV: plus ``obs_slices`` for the
Dict view).
"""
from __future__ import annotations

import numpy as np


class _Base:
    def __init__(self, env, dict_obs: bool | None = None):
        self.env = env
        self.num_envs = env.num_envs
        self.single_observation_space = env.single_observation_space
        self.single_action_space = env.single_action_space
        self.observation_space = env.single_observation_space
        self.action_space = env.single_action_space
        self._slices = getattr(env, "obs_slices", None)
        self._dict = (self._slices is not None) if dict_obs is None else dict_obs
        self._episode = 0

    def _fmt(self, flat: np.ndarray):
        if self._dict and self._slices is not None:
            return {k: flat[:, s].copy() for k, s in self._slices.items()}
        return flat

    def _fmt_one(self, flat_row: np.ndarray):
        if self._dict and self._slices is not None:
            return {k: flat_row[s].copy() for k, s in self._slices.items()}
        return flat_row.copy()

    def _reset_all(self, seed=None) -> np.ndarray:
        obs, _ = self.env.reset(seed=seed)
        self._episode += 1
        return obs.cpu().numpy() if hasattr(obs, "cpu") else np.asarray(obs)

    def _reset_done(self, done: np.ndarray, seed) -> np.ndarray:
        """Resets only the finished envs; returns the new full observation batch."""
        import inspect
        kw = {}
        if "mask" in inspect.signature(self.env.reset).parameters:
            kw["mask"] = done
        obs, _ = self.env.reset(seed=seed, **kw)
        self._episode += 1
        return obs.cpu().numpy() if hasattr(obs, "cpu") else np.asarray(obs)

    def close(self):
        self.env.close()


_NO_INFO: dict = {}       # shared by every env that has nothing to report this step (treat as read-only)


class LazyInfos:
    """The ``infos`` sequence of a vector step without per-env allocation: ``infos[i]`` is the shared empty dict unless
    env i finished this step, in which case the dict (``terminal_observation``, ``TimeLimit.truncated``) is built —
    and the terminal observations fetched from the device — on first access.  Behaves like the list SB3 expects
    (``len``, indexing, iteration); with 65 536 envs ending their day on the same step, eager construction costs
    ~0.3 s of Python per episode end."""

    def __init__(self, n: int):
        self._n = n
        self._done_idx = None
        self._fetch = None
        self._rows = None
        self._fmt_one = None
        self._cache: dict = {}
        self._pos = None

    def _set(self, done_idx, fetch, fmt_one):
        self._done_idx, self._fetch, self._fmt_one, self._rows, self._pos = done_idx, fetch, fmt_one, None, None
        self._cache = {}

    def _clear(self):
        if self._done_idx is not None:
            self._set(None, None, None)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if self._done_idx is None:
            return _NO_INFO
        if self._pos is None:
            self._pos = {int(b): j for j, b in enumerate(self._done_idx)}
        j = self._pos.get(int(i))
        if j is None:
            return _NO_INFO
        if i not in self._cache:
            if self._rows is None:
                self._rows = self._fetch()
            self._cache[i] = {"terminal_observation": self._fmt_one(self._rows[j]), "TimeLimit.truncated": False}
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(self._n))


class SB3VecEnv(_Base):
    """stable_baselines3.common.vec_env.VecEnv protocol.

    With an engine that steps asynchronously and resets in-kernel (``BatchedEVChargingEnv(auto_reset=True)``),
    ``step_async`` only enqueues H2D(actions) -> step -> D2H on the env's own stream and returns; ``step_wait``
    waits for it, so two adapters stepped alternately overlap their copies (what SubprocVecEnv's worker processes
    do in the reference, examples/evcharging/train_stable_baselines.py:264-275).  ``actions`` may be a NumPy
    array (copied into the env's pinned staging buffer) or a pinned torch tensor (DMA straight from it).
    ``infos`` is ONE list reused between steps whose entries are the shared empty dict except for envs that
    finished (``terminal_observation``, ``TimeLimit.truncated``) — no per-env allocation per step.
    ``copy=False`` returns views of the pinned result buffers (valid until the next ``step_wait``)."""

    def __init__(self, env, dict_obs=None, seed: int = 0, copy: bool = True):
        super().__init__(env, dict_obs)
        self._seed = seed
        self._actions = None
        self._copy = copy
        self._fast = bool(getattr(env, "auto_reset", False)) and hasattr(env, "step_async") and hasattr(env, "final_obs")
        self._infos = LazyInfos(self.num_envs) if self._fast else [_NO_INFO] * self.num_envs
        self._dirty = ()

    def seed(self, seed=None):
        self._seed = 0 if seed is None else seed
        return [self._seed + i for i in range(self.num_envs)]

    def reset(self):
        return self._fmt(self._reset_all(self._seed))

    def _fmt(self, flat: np.ndarray):
        if self._dict and self._slices is not None:
            return {k: (flat[:, s].copy() if self._copy else flat[:, s]) for k, s in self._slices.items()}
        return flat.copy() if self._copy else flat

    def step_async(self, actions):
        if self._fast:
            env = self.env
            if isinstance(actions, np.ndarray):
                env._h_action.numpy()[...] = actions.reshape(env._h_action.shape)
                actions = env._h_action
            env.step_async(actions)
        else:
            self._actions = np.asarray(actions)

    def step_wait(self):
        infos = self._infos
        if self._fast:
            infos._clear()
            obs, rew, term, trunc, _ = self.env.step_wait()
            obs, done = obs.numpy(), term.numpy()
            if done.any():                                 # the kernel already reset them: obs holds the reset observation
                idx = np.nonzero(done)[0]
                env = self.env
                final = env.final_obs.clone()              # (device copy: the next step may overwrite final_obs)

                def fetch(idx=idx, final=final, env=env):
                    import torch
                    return final[torch.from_numpy(idx).to(env.device)].cpu().numpy()
                infos._set(idx, fetch, self._fmt_one)
                self._episode += 1
            rew = rew.numpy().astype(np.float32)
            return self._fmt(obs), rew, (done.copy() if self._copy else done), infos
        for i in self._dirty:                              # entries filled by the previous step
            infos[i] = _NO_INFO
        self._dirty = ()
        obs, rew, term, trunc, _ = self.env.step(self._actions)
        obs = np.asarray(obs)
        done = np.asarray(term) | np.asarray(trunc)
        if done.any():
            terminal = obs.copy()
            obs = self._reset_done(done, self._seed + self._episode * self.num_envs)
            idx = np.nonzero(done)[0]
            for i in idx:
                infos[i] = {"terminal_observation": self._fmt_one(terminal[i]),
                            "TimeLimit.truncated": bool(np.asarray(trunc)[i] and not np.asarray(term)[i])}
            self._dirty = idx
        return self._fmt(obs), np.asarray(rew, np.float32), done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def get_attr(self, name, indices=None):
        return [getattr(self.env, name)] * self.num_envs


class GymnasiumVectorEnv(_Base):
    """gymnasium.vector.VectorEnv protocol (0.28: final_observation / final_info on auto-reset)."""

    def reset(self, *, seed=None, options=None):
        return self._fmt(self._reset_all(seed)), {}

    def step(self, actions):
        obs, rew, term, trunc, _ = self.env.step(np.asarray(actions))
        obs, term, trunc = np.asarray(obs), np.asarray(term), np.asarray(trunc)
        done = term | trunc
        infos = {}
        if done.any():
            final = np.empty(self.num_envs, dtype=object)
            final_info = np.empty(self.num_envs, dtype=object)
            terminal = obs.copy()
            obs = self._reset_done(done, self._episode * self.num_envs)
            for i in np.nonzero(done)[0]:
                final[i] = self._fmt_one(terminal[i])
                final_info[i] = {}
            infos = {"final_observation": final, "_final_observation": done.copy(),
                     "final_info": final_info, "_final_info": done.copy()}
        return self._fmt(obs), np.asarray(rew, np.float64), term, trunc, infos


class RLlibVectorEnv(_Base):
    """ray.rllib.env.vector_env.VectorEnv protocol."""

    def vector_reset(self, *, seeds=None, options=None):
        obs = self._reset_all(None if seeds is None else np.asarray(seeds))
        self._last = obs
        return [self._fmt_one(o) for o in obs], [{} for _ in range(self.num_envs)]

    def reset_at(self, index=None, *, seed=None, options=None):
        mask = np.zeros(self.num_envs, bool)
        mask[0 if index is None else index] = True
        obs = self._reset_done(mask, self._episode if seed is None else seed)
        self._last = obs
        return self._fmt_one(obs[0 if index is None else index]), {}

    def vector_step(self, actions):
        obs, rew, term, trunc, _ = self.env.step(np.asarray(actions))
        obs = np.asarray(obs)
        self._last = obs
        return ([self._fmt_one(o) for o in obs], [float(r) for r in rew], [bool(t) for t in term],
                [bool(t) for t in trunc], [{} for _ in range(self.num_envs)])

    def get_sub_environments(self):
        return []
