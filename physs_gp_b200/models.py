"""`SDE_GP`: the call site of the hot path -- mirror of stgp/models/sde_gp.py:155-302
(log_marginal_likelihood, filter, filter_and_smooth, posterior_blocks, posterior)."""
import numpy as np
import torch

from . import filters
from .data import TemporalData
from .likelihood import get_R_R_inv


class SDE_GP:
    def __init__(self, data, prior, likelihood, full_state_observed=False, filter_type='b200'):
        self.data = data
        self.prior = prior
        self.likelihood = likelihood
        self.full_state_observed = full_state_observed
        self.filter_type = filter_type

    def _R(self):
        return get_R_R_inv(self.likelihood, self.data.Nt, self.data.P * self.data.Ns)

    def log_marginal_likelihood(self):
        R, R_inv = self._R()
        lml, _ = filters.filter_loop(self.data, self.prior, R=R, R_inv=R_inv, filter_type=self.filter_type)
        return lml

    def get_objective(self):
        return -self.log_marginal_likelihood()

    def filter(self, return_lml=False):
        R, R_inv = self._R()
        lml, kf = filters.filter_loop(self.data, self.prior, R=R, R_inv=R_inv, filter_type=self.filter_type)
        return (lml, kf['m'], kf['P']) if return_lml else (kf['m'], kf['P'])

    def filter_and_smooth(self, full_state=False, return_lml=False):
        R, R_inv = self._R()
        lml, kf = filters.filter_loop(self.data, self.prior, R=R, R_inv=R_inv, filter_type=self.filter_type)
        mu, var = filters.smoother_loop(self.data, self.prior, kf, full_state=full_state,
                                        filter_type=self.filter_type)
        return (lml, mu, var) if return_lml else (mu, var)

    def posterior_blocks(self, return_lml=False):
        """sde_gp.py:255-277: mu [T, m', 1], var [T, 1, m', m'] (leading B when batched)."""
        lml, mu, var = self.filter_and_smooth(return_lml=True)
        var = var.unsqueeze(-3)
        return (lml, mu, var) if return_lml else (mu, var)

    def posterior(self, diagonal=True, full_state=False):
        mu, var = self.filter_and_smooth(full_state=full_state)
        if not full_state and diagonal:
            return mu, var.diagonal(dim1=-2, dim2=-1)[..., None]
        return mu, var

    # ---------------------------------------------------------------- prediction at new times
    def predict_f(self, XS, diagonal=True, squeeze=False, filter_only=False, force_full_state=False):
        """Posterior at new time points -- mirror of `predict_f` (stgp/models/sde_gp.py:392-488): the test
        times are stacked BEHIND the training times (so that a test time equal to a training time keeps the
        training row), sorted and de-duplicated (`order_sequentially_np`, stgp/data/sequential.py:85-144:
        `np.unique(..., return_index, return_inverse)`), run through the same filter / smoother with NaN
        observations at the test rows, then unsorted and the training rows dropped
        (`SequentialData.unsort`, stgp/data/data.py:413-415).

        XS [NS] or [NS, 1].  Returns (mu [NS, out, 1], var [NS, out, 1, 1]) with `diagonal`, var
        [NS, 1, out, out] without; out = output dim, or the state dim with `filter_only` /
        `force_full_state`; a leading batch axis B when the data is batched (shared time grid)."""
        if filter_only and force_full_state:
            raise RuntimeWarning('filter_only alrady returns the full state')
        data = self.data
        X = data.X_time.detach().cpu().numpy() if isinstance(data.X_time, torch.Tensor) else np.asarray(data.X_time)
        X = X.reshape(-1)
        XS = XS.detach().cpu().numpy() if isinstance(XS, torch.Tensor) else np.asarray(XS, dtype=np.float64)
        XS = XS.reshape(-1)
        NS, Nt = XS.shape[0], X.shape[0]
        stacked = np.concatenate([X, XS])
        X_sorted, unique_idx, reverse_idx = np.unique(stacked, return_index=True, return_inverse=True)
        Y = data.Y_st
        if isinstance(Y, torch.Tensor):
            pad = torch.full((*Y.shape[:-3], NS, *Y.shape[-2:]), float('nan'), dtype=Y.dtype, device=Y.device)
            Y_sorted = torch.cat([Y, pad], dim=-3)[..., torch.as_tensor(unique_idx, device=Y.device), :, :]
        else:
            Y = np.asarray(Y)
            pad = np.full((*Y.shape[:-3], NS, *Y.shape[-2:]), np.nan)
            Y_sorted = np.concatenate([Y, pad], axis=-3)[..., unique_idx, :, :]
        test_data = TemporalData(X_sorted, Y_sorted, data.X_space)
        m = data.P * data.Ns
        R = self.likelihood.R_predict(Nt, NS, unique_idx, m)
        lml, kf = filters.filter_loop(test_data, self.prior, R=R, R_inv=None, filter_type=self.filter_type)
        if filter_only:
            mu, var = kf['m'], kf['P']
        else:
            mu, var = filters.smoother_loop(test_data, self.prior, kf, full_state=force_full_state,
                                            filter_type=self.filter_type)
        idx = torch.as_tensor(np.asarray(reverse_idx).reshape(-1)[Nt:], device=mu.device)
        mu = mu.index_select(-3, idx)                      # [.., NS, out, 1]
        var = var.index_select(-3, idx)                    # [.., NS, out, out]
        if diagonal:
            var = var.diagonal(dim1=-2, dim2=-1)           # [.., NS, out]
        if squeeze:
            return mu.squeeze(), var.squeeze()
        if diagonal:
            return mu, var[..., None, None]
        return mu, var.unsqueeze(-3)
