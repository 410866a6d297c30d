"""``DiscreteActionWrapper`` — mirrors ``sustaingym/envs/wrappers.py``.

For the batched EV env the conversion ``{0..bins-1} -> k/(bins-1)`` is fused into the step kernel
(``discrete_bins`` constructor argument, ``SG_EV_ACTION_I64``); this class wraps any single-env view
with a Box action space for API parity."""
from __future__ import annotations

import numpy as np

from .. import spaces


class DiscreteActionWrapper:
    def __init__(self, env, bins: int = 5):
        if not isinstance(env.action_space, spaces.Box):
            raise ValueError("Should only be used to wrap continuous env")
        self.env = env
        self._bins = bins
        self._cont_dtype = env.action_space.dtype
        dims = env.action_space.shape
        self.action_space = spaces.Discrete(bins) if len(dims) == 0 else \
            spaces.MultiDiscrete(np.ones(dims, dtype=np.int64) * bins)
        self.observation_space = env.observation_space

    def __repr__(self) -> str:
        return repr(self.env)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def action(self, action):
        """wrappers.py:43-45."""
        return np.asarray(action, dtype=self._cont_dtype) / (self._bins - 1)

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(self.action(action))
