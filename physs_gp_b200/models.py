"""`SDE_GP`: the call site of the hot path -- mirror of stgp/models/sde_gp.py:155-302
(log_marginal_likelihood, filter, filter_and_smooth, posterior_blocks, posterior)."""
from . import filters
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
