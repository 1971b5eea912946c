"""Minimal Box / MultiDiscrete spaces with the attributes agents read (`shape`, `low`, `high`, `nvec`,
`dtype`, `sample()`, `seed()`, `contains()`).  The image has no `gym`; if a real `gym` is importable its
own space classes are used instead so that `isinstance(space, gym.spaces.Box)` holds for agent code.
Reference: spaces are declared at supplychain_env.py:625-626 and beergame2_env.py:27-28."""
import numpy as np

try:  # pragma: no cover - gym is not installed in the build image
    from gym.spaces import Box, MultiDiscrete  # noqa: F401
except Exception:  # noqa: BLE001
    class _Space(object):
        def __init__(self, shape, dtype):
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def __contains__(self, x):
            return self.contains(x)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            shape = np.shape(low) if shape is None else tuple(shape)
            super().__init__(shape, dtype)
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)

        def sample(self):
            return self._rng.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

    class MultiDiscrete(_Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            super().__init__(self.nvec.shape, np.int64)

        def sample(self):
            return (self._rng.random_sample(self.nvec.shape) * self.nvec).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0)) and bool(np.all(x < self.nvec))

        def __repr__(self):
            return "MultiDiscrete(%s)" % (self.nvec,)
