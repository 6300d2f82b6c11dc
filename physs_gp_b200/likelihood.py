"""Likelihood containers the path reads R from -- mirror of stgp/likelihood/gaussian.py and
`get_R_R_inv` (stgp/models/sde_gp.py:30-43)."""
import numpy as np


class Gaussian:
    """Homoscedastic Gaussian noise, R_k = variance * I_m for every step."""

    def __init__(self, variance=1.0):
        self.variance_scalar = float(variance)

    def R(self, Nt, m):
        return self.variance_scalar * np.eye(m)[None]          # [1, m, m], broadcast over time

    def R_predict(self, Nt, NS, unique_idx, m):
        """Noise for the merged train + test grid of `predict_f` (sde_gp.py get_likelihood_for_prediction)."""
        return self.R(len(unique_idx), m)


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

    def R_predict(self, Nt, NS, unique_idx, m):
        """Per-step blocks on the merged grid: the training blocks where a training row survives, identity at
        the test rows (their observations are NaN: the block only enters the masked innovation covariance)."""
        import torch
        V = self._variance
        if isinstance(V, torch.Tensor):
            eye = torch.eye(m, dtype=V.dtype, device=V.device).expand(*V.shape[:-3], NS, m, m)
            return torch.cat([V, eye], dim=-3)[..., torch.as_tensor(unique_idx, device=V.device), :, :]
        V = np.asarray(V)
        eye = np.broadcast_to(np.eye(m), (*V.shape[:-3], NS, m, m))
        return np.concatenate([V, eye], axis=-3)[..., unique_idx, :, :]


class PrecisionBlockDiagonalGaussian(BlockDiagonalGaussian):
    """BlockDiagonalGaussian storing the PRECISION blocks [.., Nt, m, m] (likelihood/gaussian.py:96-105): what the
    'NG_Precision' CVI sites live in.  `variance` is the reference's `mat_inv` of the precision
    (computation/matrix_ops.py:383-385: Cholesky of precision + settings.jitter I, then solve on the identity)."""

    def __init__(self, precision):
        self._precision = precision

    @property
    def precision(self):
        return self._precision

    @property
    def variance(self):
        import torch
        from . import ops, settings
        P = self._precision
        if not isinstance(P, torch.Tensor):
            P = torch.as_tensor(np.asarray(P, np.float64)).cuda()
        return ops.spd_inverse(P, settings.jitter)

    def R(self, Nt, m):
        return self.variance

    def R_predict(self, Nt, NS, unique_idx, m):
        return BlockDiagonalGaussian(self.variance).R_predict(Nt, NS, unique_idx, m)


def get_R_R_inv(likelihood, Nt, m):
    """sde_gp.py:30-43: (variance, None), or (None, precision) for a PrecisionBlockDiagonalGaussian."""
    if isinstance(likelihood, PrecisionBlockDiagonalGaussian):
        return None, likelihood.precision
    return likelihood.R(Nt, m), None
