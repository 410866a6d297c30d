"""Minimal Gymnasium-compatible space objects.

The reference declares its spaces with ``gymnasium.spaces`` (gymnasium 0.28).  gymnasium is not
installed in this image, so the engine ships duck-typed equivalents with the attributes RL
libraries read (``shape``, ``dtype``, ``low``, ``high``, ``n``, ``nvec``, ``spaces``, ``sample``,
``contains``, ``seed``).  When gymnasium IS importable the same classes are taken from it, so the
batched envs expose genuine gymnasium spaces to StableBaselines3 / RLlib.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - not available in the build image
    from gymnasium.spaces import Box, Dict, Discrete, MultiDiscrete, flatten, flatten_space  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Space:
        shape: tuple = ()
        dtype = None

        def __init__(self, seed=None):
            self._np_random = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._np_random = np.random.default_rng(seed)
            return [seed]

        @property
        def np_random(self):
            return self._np_random

        def __contains__(self, x):
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            super().__init__(seed)
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._np_random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return (x.shape == self.shape and np.can_cast(x.dtype, self.dtype, casting="same_kind")
                    and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high)))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, o):
            return isinstance(o, Box) and self.shape == o.shape and self.dtype == o.dtype \
                and np.array_equal(self.low, o.low) and np.array_equal(self.high, o.high)

    class Discrete(Space):
        def __init__(self, n, start=0, seed=None):
            super().__init__(seed)
            self.n, self.start = int(n), int(start)
            self.shape, self.dtype = (), np.dtype(np.int64)

        def sample(self):
            return np.int64(self.start + self._np_random.integers(self.n))

        def contains(self, x):
            try:
                xi = int(x)
            except (TypeError, ValueError):
                return False
            return xi == x and self.start <= xi < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n}, start={self.start})" if self.start else f"Discrete({self.n})"

        def __eq__(self, o):
            return isinstance(o, Discrete) and (self.n, self.start) == (o.n, o.start)

    class MultiDiscrete(Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            super().__init__(seed)
            self.nvec = np.asarray(nvec, dtype=dtype)
            self.shape, self.dtype = self.nvec.shape, np.dtype(dtype)

        def sample(self):
            return (self._np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and np.issubdtype(x.dtype, np.integer) \
                and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))

        def __getitem__(self, i):
            return Discrete(int(self.nvec[i]))

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Dict(Space):
        """Keys are kept sorted, as gymnasium 0.28 does for plain-dict input."""

        def __init__(self, spaces=None, seed=None):
            super().__init__(seed)
            spaces = dict(spaces or {})
            self.spaces = OrderedDict((k, spaces[k]) for k in sorted(spaces))

        def seed(self, seed=None):
            super().seed(seed)
            ss = np.random.SeedSequence(seed).spawn(len(self.spaces))
            for sp, sq in zip(self.spaces.values(), ss):
                sp.seed(int(sq.generate_state(1)[0]))
            return [seed]

        def sample(self):
            return OrderedDict((k, sp.sample()) for k, sp in self.spaces.items())

        def contains(self, x):
            return isinstance(x, dict) and set(x) == set(self.spaces) \
                and all(sp.contains(x[k]) for k, sp in self.spaces.items())

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {v!r}" for k, v in self.spaces.items()) + ")"

    def flatten_space(space):
        """gymnasium.spaces.flatten_space for the space kinds the reference uses."""
        if isinstance(space, Box):
            return Box(space.low.ravel(), space.high.ravel(), dtype=space.dtype)
        if isinstance(space, Discrete):
            return Box(0, 1, (space.n,), dtype=np.int64)
        if isinstance(space, MultiDiscrete):
            return Box(0, 1, (int(space.nvec.sum()),), dtype=np.int64)
        if isinstance(space, Dict):
            parts = [flatten_space(s) for s in space.spaces.values()]
            dt = np.result_type(*[p.dtype for p in parts])
            return Box(np.concatenate([p.low for p in parts]).astype(dt),
                       np.concatenate([p.high for p in parts]).astype(dt), dtype=dt)
        raise NotImplementedError(type(space))

    def flatten(space, x):
        """gymnasium.spaces.flatten: Box -> ravel, Discrete -> one-hot, Dict -> concat in key order."""
        if isinstance(space, Box):
            return np.asarray(x, dtype=space.dtype).ravel()
        if isinstance(space, Discrete):
            out = np.zeros(space.n, np.int64)
            out[int(x) - space.start] = 1
            return out
        if isinstance(space, MultiDiscrete):
            out = np.zeros(int(space.nvec.sum()), np.int64)
            off = np.concatenate([[0], np.cumsum(space.nvec)[:-1]])
            out[off + np.asarray(x)] = 1
            return out
        if isinstance(space, Dict):
            return np.concatenate([np.asarray(flatten(s, x[k]), dtype=flatten_space(space).dtype)
                                   for k, s in space.spaces.items()])
        raise NotImplementedError(type(space))
