"""gym.utils.seeding.np_random as of gym 0.10.9: returns (RandomState, seed)."""
import numpy as np


def np_random(seed=None):
    if seed is None:
        seed = int(np.random.SeedSequence().entropy % (2 ** 31))
    rng = np.random.RandomState(seed % (2 ** 32))
    return rng, seed
