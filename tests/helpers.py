"""Shared seeded synthetic inputs for the parity tests (same generators on CPU and GPU side)."""
import numpy as np


def myexp2(x1, x2):
    """code/test_solvers.py:6-11 -- the reference's own test function."""
    z = 2 * np.maximum(0, x1 + x2)
    return np.exp(z) - (z + z ** 2 / 2 + z ** 3 / 6)


def synth(seed, n, p, lo=-1.0, hi=1.0, sigma=1.0):
    """Noisy function on U(lo,hi)^p; p==2 uses the reference's myexp2."""
    rng = np.random.RandomState(seed)
    x = rng.uniform(lo, hi, (n, p))
    if p == 2:
        f = myexp2(x[:, 0], x[:, 1])
    else:
        f = np.sin(3 * x[:, 0])
        for k in range(1, p):
            f = f + (k + 1) * (x[:, k] > 0.2 * (hi + lo + 1))
    return x, f + sigma * rng.normal(size=n)
