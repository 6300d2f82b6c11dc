"""Likelihood containers the path reads R from -- mirror of stgp/likelihood/gaussian.py and
`get_R_R_inv` (stgp/models/sde_gp.py:30-43)."""
import numpy as np


class Gaussian:
    """Homoscedastic Gaussian noise, R_k = variance * I_m for every step."""

    def __init__(self, variance=1.0):
        self.variance_scalar = float(variance)

    def R(self, Nt, m):
        return self.variance_scalar * np.eye(m)[None]          # [1, m, m], broadcast over time


class BlockDiagonalGaussian:
    """Per-step full m x m noise blocks [Nt, m, m] (the CVI site covariance V-tilde lives here,
    likelihood/gaussian.py:35-93)."""

    def __init__(self, variance):
        self._variance = variance

    @property
    def variance(self):
        return self._variance

    def R(self, Nt, m):
        return self._variance


def get_R_R_inv(likelihood, Nt, m):
    return likelihood.R(Nt, m), None
