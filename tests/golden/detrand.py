"""Version-independent deterministic test data (splitmix64 counter hash in numpy), shared by
``make_golden.py`` and the tests so that fixtures store seeds + reference OUTPUTS only."""
import numpy as np


def _mix(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def bits(seed: int, shape) -> np.ndarray:
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3)
        return _mix(_mix(idx)).reshape(shape)


def uniform(seed: int, shape) -> np.ndarray:
    """float32 uniforms in [0, 1) with 24 random bits (like torch.rand)."""
    return ((bits(seed, shape) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24))


def integers(seed: int, lo: int, hi: int, shape) -> np.ndarray:
    return (lo + (bits(seed, shape) >> np.uint64(11)) % np.uint64(hi - lo)).astype(np.int64)


def normal(seed: int, shape) -> np.ndarray:
    """float32 standard normals (Box-Muller on two uniform streams)."""
    u1 = np.maximum(uniform(seed, shape).astype(np.float64), 2.0 ** -24)
    u2 = uniform(seed + 7919, shape).astype(np.float64)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)).astype(np.float32)
