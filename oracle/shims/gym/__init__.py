"""Stub of the slice of gym 0.10.9 the reference imports (ship_env.py:6-8).  TEST INFRASTRUCTURE ONLY."""
from . import spaces, utils  # noqa: F401


class Env(object):
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        return

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self
