"""Minimal ``gym.spaces`` stand-ins (the reference imports gym only for these; environment.py:1-3,65-96).

Only shape / dtype / bounds / ``sample()`` are used by gym-formation and its trainers, so the drop-in
carries its own tiny classes instead of depending on gym (which current Pythons cannot install)."""
import numpy as np


class Space(object):
    shape = ()
    dtype = None


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        # sampling bounds (unbounded sides fall back to [-1, 1]); computed once: sample() is on the per-step path of
        # the reference's demo loop (test.py:20)
        self._lo = np.where(np.isfinite(self.low), self.low, -1.0).astype(np.float64)
        self._hi = np.where(np.isfinite(self.high), self.high, 1.0).astype(np.float64)

    def sample(self):
        return np.random.uniform(self._lo, self._hi, self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)
