"""Seeding of numpy's global generator.

The engine keeps parameter initialisation and batch shuffling on the host precisely so that
`random_seed(s)` reproduces the reference's weights and batch order bit for bit."""
import numpy as np

_SEED_LIMIT = 1 << 32


def random_seed(seed):
    """Seed np.random; `seed` must fit an unsigned 32-bit integer (ValueError otherwise)."""
    value = int(seed)
    if value not in range(_SEED_LIMIT):
        raise ValueError("Seed must be between 0 and 2**32 - 1")
    np.random.seed(value)
    return value
