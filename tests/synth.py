"""Seeded synthetic inputs of the BASELINE.json shapes, shared by the parity
tests, smoke() and bench.py.  Pure numpy; nothing here touches the product or the oracle."""
import numpy as np


def time_grid(T, dt0=0.1, rng=None, irregular=True):
    rng = rng or np.random.default_rng(0)
    steps = rng.uniform(0.5, 1.5, T) * dt0 if irregular else np.full(T, dt0)
    return np.cumsum(steps)


def noisy_series(B, T, m, rng, nan_frac=0.05, scale=1.0):
    """Smooth-ish signal + noise with `nan_frac` entries missing.  [B, T, m]"""
    t = np.arange(T)[None, :, None]
    phase = rng.uniform(0, 2 * np.pi, (B, 1, m))
    freq = rng.uniform(0.01, 0.05, (B, 1, m))
    Y = scale * np.sin(freq * t + phase) + 0.3 * rng.normal(size=(B, T, m))
    if nan_frac > 0:
        Y[rng.uniform(size=Y.shape) < nan_frac] = np.nan
    return Y


def random_spd(rng, shape_prefix, m, base=0.1, spread=0.2):
    G = rng.normal(size=tuple(shape_prefix) + (m, m)) * spread
    return G @ np.swapaxes(G, -1, -2) + base * np.eye(m)


def log_uniform(rng, lo, hi, size):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size))
