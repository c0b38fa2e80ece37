"""Minimal stand-ins for the parts of gym 0.17 the reference API exposes
(`gym.Env`, `spaces.Discrete/Box/Dict`); used because gym is not installed in
this image.  If gym is importable its own classes are used instead."""
import collections

import numpy as np

try:  # pragma: no cover - gym is absent in the build image
    import gym as _gym
    from gym import spaces as _spaces
    Env = _gym.Env
    Discrete, Box, Dict = _spaces.Discrete, _spaces.Box, _spaces.Dict
    HAVE_GYM = True
except Exception:  # noqa: BLE001
    HAVE_GYM = False

    class Space:
        def __init__(self):
            self.np_random = np.random.RandomState()

        def seed(self, seed=None):
            self.np_random = np.random.RandomState(seed)
            return [seed]

    class Discrete(Space):
        def __init__(self, n):
            super().__init__()
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def sample(self):
            return int(self.np_random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return f'Discrete({self.n})'

        def __eq__(self, other):
            return isinstance(other, Discrete) and other.n == self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            super().__init__()
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype),
                                       self.shape)
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype),
                                        self.shape)

        def sample(self):
            return self.np_random.randint(
                0, 256, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and np.all(x >= self.low) \
                and np.all(x <= self.high)

        def __repr__(self):
            return f'Box{self.shape}'

    class Dict(Space):
        def __init__(self, spaces):
            super().__init__()
            self.spaces = collections.OrderedDict(spaces)

        def sample(self):
            return collections.OrderedDict(
                (k, s.sample()) for k, s in self.spaces.items())

        def contains(self, x):
            return all(k in x and s.contains(x[k])
                       for k, s in self.spaces.items())

        def __repr__(self):
            return f'Dict({dict(self.spaces)})'

    class Env:
        metadata = {'render.modes': ['rgb_array']}
        reward_range = (-float('inf'), float('inf'))
        action_space = None
        observation_space = None

        @property
        def unwrapped(self):
            return self

        def __enter__(self):
            return self

        def __exit__(self, *args):
            self.close()
            return False
