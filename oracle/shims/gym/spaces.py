"""gym.spaces.Box / Discrete as of gym 0.10.9 (only what ship_env.py:19,48,143 needs)."""
import numpy as np

np_random = np.random.RandomState(0)


class Discrete(object):
    def __init__(self, n):
        self.n = n
        self.shape = ()
        self.dtype = np.int64

    def sample(self):
        return np_random.randint(self.n)

    def contains(self, x):
        if isinstance(x, int):
            as_int = x
        elif isinstance(x, (np.generic, np.ndarray)) and (x.dtype.kind in np.typecodes["AllInteger"] and x.shape == ()):
            as_int = int(x)
        else:
            return False
        return 0 <= as_int < self.n


class Box(object):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low)
        self.high = np.full(self.shape, high)
