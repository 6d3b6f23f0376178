import numpy as np


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s)" % (self.shape, self.dtype, "...")


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n, self.start, self.shape, self.dtype = int(n), int(start), (), np.dtype(np.int64)


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec); self.shape = self.nvec.shape; self.dtype = np.dtype(np.int64)


class MultiBinary(Space):
    def __init__(self, n):
        self.n = n; self.shape = (n,) if isinstance(n, int) else tuple(n); self.dtype = np.dtype(np.int8)


class Dict(Space):
    def __init__(self, spaces=None):
        self.spaces = dict(spaces or {})


class Tuple(Space):
    def __init__(self, spaces=()):
        self.spaces = tuple(spaces)


class utils:          # gymnasium.spaces.utils (stable_baselines3/common/preprocessing.py:186)
    @staticmethod
    def flatdim(space):
        if isinstance(space, Box):
            return int(np.prod(space.shape))
        if isinstance(space, Discrete):
            return int(space.n)
        if isinstance(space, MultiDiscrete):
            return int(np.sum(space.nvec))
        if isinstance(space, MultiBinary):
            return int(np.prod(space.shape))
        raise NotImplementedError(type(space))
