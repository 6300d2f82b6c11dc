"""`SDE_GP`: the call site of the hot path -- mirror of stgp/models/sde_gp.py:155-302
(log_marginal_likelihood, filter, filter_and_smooth, posterior_blocks, posterior)."""
import numpy as np
import torch

from . import filters
from .data import TemporalData
from .likelihood import get_R_R_inv


def _block_diag_batched(blocks):
    """[..., s, s] blocks -> [..., d, d] block-diagonal (differentiable)."""
    lead = blocks[0].shape[:-2]
    sizes = [b.shape[-1] for b in blocks]
    d = sum(sizes)
    rows, o = [], 0
    for b, sz in zip(blocks, sizes):
        left = torch.zeros(lead + (sz, o), dtype=b.dtype, device=b.device)
        right = torch.zeros(lead + (sz, d - o - sz), dtype=b.dtype, device=b.device)
        rows.append(torch.cat([left, b, right], dim=-1))
        o += sz
    return torch.cat(rows, dim=-2)


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
        if not full_state:
            # nobody reads the filtered moments: one call, packed hand-over (None = outside what it covers)
            fused = filters.filter_smooth_fused(self.data, self.prior, R=R, R_inv=R_inv, filter_type=self.filter_type)
            if fused is not None:
                return fused if return_lml else fused[1:]
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
        """sde_gp.py:279-305: with `full_state=False` the mean is stacked time-space to [Nt * m', 1] and, with
        `diagonal`, the variances to [Nt * m', 1] (a leading B when batched); otherwise the block outputs."""
        mu, var = self.filter_and_smooth(full_state=full_state)
        if full_state:
            return mu, var
        lead = tuple(mu.shape[:-3])
        mu = mu.reshape(lead + (-1, 1))
        if diagonal:
            return mu, var.diagonal(dim1=-2, dim2=-1).reshape(lead + (-1, 1))
        return mu, var

    # ---------------------------------------------------------------- lml and its hyper-parameter gradient
    def log_marginal_likelihood_and_grad(self):
        """Value and gradient of the lml with respect to the kernel and noise hyper-parameters -- the pair
        the reference's trainers take with `jax.value_and_grad` / `jacrev` THROUGH the filter
        (stgp/trainers/standard.py:58-91, stgp/trainers/trainer.py:43,128-136).  One forward filter launch, one
        reverse launch (`physs_kf_filter_vjp_f64`), then the T-independent chain of the prior's closed forms on
        the host.  For `BatchedMaternSDE` priors with scalar Gaussian observations (d <= 4).

        Returns (lml [B], {'lengthscale': [B, nblk], 'variance': [B, nblk], 'noise': [B]})."""
        from . import ops, settings
        from .likelihood import Gaussian
        from .sdes import BatchedMaternSDE
        prior, data = self.prior, self.data
        if not isinstance(prior, BatchedMaternSDE):
            raise NotImplementedError("lml gradient: BatchedMaternSDE prior")
        if data.P * data.Ns != 1 or prior.full_state_obs or prior.d > 4 or not isinstance(self.likelihood, Gaussian):
            return self._lml_and_grad_general()
        dev = filters._device()
        X_t = filters._time_axis(data, dev)
        dt = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), X_t[1:] - X_t[:-1]])
        Y = filters._to_dev(data.Y_st, dev)
        Y = Y.reshape(*Y.shape[:-2], 1)                           # [.., Nt, 1]
        if Y.dim() == 2:
            Y = Y[None]
        if Y.shape[0] == 1 and prior.B > 1:
            Y = Y.expand(prior.B, -1, -1)
        if settings.time_major and Y.shape[0] >= settings.time_major_min_batch and not ops.step_layout(Y, "Y")[1]:
            Y = Y.transpose(0, 1).contiguous().transpose(0, 1)
        (disc,), m0, P0, H = filters.lower_prior(prior, data.X_space, [dt], dev)
        R = torch.full((1, 1, 1, 1), self.likelihood.variance_scalar, dtype=torch.float64, device=dev)
        Hd = filters._to_dev(H, dev)
        lml, mf, Pf = ops.kf_filter(dt, Y, R, Hd, m0, P0, disc, jitter=settings.jitter)
        g = ops.kf_filter_vjp(dt, Y, R, Hd, m0, P0, disc, mf, Pf, jitter=settings.jitter)
        g_ls, g_var = prior.hyper_grads(g['glam'].cpu().numpy(), (g['gPinf'] + g['gP0']).cpu().numpy())
        return lml, {'lengthscale': torch.as_tensor(g_ls, device=dev), 'variance': torch.as_tensor(g_var, device=dev),
                     'noise': g['gR'][:, 0, 0]}

    def _lml_and_grad_general(self):
        """The same pair for any state / observation dimension (d <= 32, m <= d; e.g. full-state sites m = d at
        d = 6 .. 12, the shapes a VB_NG_ADAM epoch differentiates through): the transitions A_k = expm(F dt_k),
        Q_k = Pinf - A_k Pinf A_k^T (kernels/kernel.py:207-209) are built ONCE with torch from the kernel
        hyper-parameters, the filter and its reverse pass run on them (`physs_kf_filter_vjp_f64`, lane-group
        kernel: gA_k, gQ_k, gP0, gR), and torch differentiates the T-independent closed forms.  Materialises
        [B, T, d, d] transitions: meant for few series.  Likelihood: `Gaussian` (adds 'noise') or
        `BlockDiagonalGaussian` site covariances (kernel gradients only)."""
        from . import ops, settings
        from .likelihood import Gaussian
        prior, data = self.prior, self.data
        dev = filters._device()
        X_t = filters._time_axis(data, dev)
        dt = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), X_t[1:] - X_t[:-1]])
        Y = filters._to_dev(data.Y_st, dev)
        Y = Y.reshape(*Y.shape[:-2], -1)                          # [.., Nt, m]
        if Y.dim() == 2:
            Y = Y[None]
        B, T, m = prior.B, Y.shape[1], Y.shape[2]
        if Y.shape[0] == 1 and B > 1:
            Y = Y.expand(B, -1, -1).contiguous()
        s, nblk, d = prior.s, prior.nblk, prior.d
        ls = torch.tensor(prior.ls, dtype=torch.float64, device=dev, requires_grad=True)
        var = torch.tensor(prior.var, dtype=torch.float64, device=dev, requires_grad=True)
        lam = (2.0 * s - 1.0) ** 0.5 / ls                          # [B, nblk]
        # companion drift of a Matern-(s - 1/2) block: last row -C(s, i) lam^(s - i)
        import math
        F = torch.zeros((B, nblk, s, s), dtype=torch.float64, device=dev)
        for i in range(s - 1):
            F[..., i, i + 1] = 1.0
        rows = [-(math.comb(s, i)) * lam ** (s - i) for i in range(s)]
        F = F + torch.stack([torch.zeros_like(lam)] * (s * (s - 1)) + rows, dim=-1).reshape(B, nblk, s, s)
        udt, inv = torch.unique(dt, return_inverse=True)
        E = torch.linalg.matrix_exp(F[:, :, None] * udt[None, None, :, None, None])          # [B, nblk, U, s, s]
        Ab = E[:, :, inv]                                                                    # [B, nblk, T, s, s]
        # stationary covariance blocks (closed forms of kernels.py / BatchedMaternSDE.P_inf)
        z = torch.zeros_like(lam)
        if s == 1:
            Pb = var[..., None, None]
        elif s == 2:
            Pb = torch.stack([var, z, z, lam ** 2 * var], -1).reshape(B, nblk, 2, 2)
        elif s == 3:
            k = lam ** 2 * var / 3.0
            Pb = torch.stack([var, z, -k, z, k, z, -k, z, lam ** 4 * var], -1).reshape(B, nblk, 3, 3)
        else:
            k1, k2 = lam ** 2 * var / 5.0, lam ** 4 * var / 5.0
            Pb = torch.stack([var, z, -k1, z, z, k1, z, -k2, -k1, z, k2, z, z, -k2, z, lam ** 6 * var],
                             -1).reshape(B, nblk, 4, 4)
        A = _block_diag_batched([Ab[:, q] for q in range(nblk)])                              # [B, T, d, d]
        Pinf = _block_diag_batched([Pb[:, q] for q in range(nblk)])                           # [B, d, d]
        Q = Pinf[:, None] - A @ Pinf[:, None] @ A.transpose(-1, -2)
        H = prior.H()
        Hd = None if (H.shape[0] == H.shape[1] and np.array_equal(H, np.eye(H.shape[0]))) else filters._to_dev(H, dev)
        if isinstance(self.likelihood, Gaussian):
            R = (self.likelihood.variance_scalar * torch.eye(m, dtype=torch.float64, device=dev)).reshape(1, 1, m, m)
        else:
            R = filters._to_dev(self.likelihood.variance, dev)
            R = R if R.dim() == 4 else R[None]
        R = R.expand(B, T, m, m).contiguous()
        dtb = dt[None].expand(B, T).contiguous()
        m0 = torch.zeros((B, d), dtype=torch.float64, device=dev)
        Ad, Qd, P0 = A.detach().contiguous(), Q.detach().contiguous(), Pinf.detach().contiguous()
        disc = ops.Disc.given(Ad, Qd)
        lml, mf, Pf = ops.kf_filter(dtb, Y, R, Hd, m0, P0, disc, jitter=settings.jitter)
        g = ops.kf_filter_vjp(dtb, Y, R, Hd, m0, P0, disc, mf, Pf, jitter=settings.jitter)
        surrogate = (g['gA'] * A).sum() + (g['gQ'] * Q).sum() + (g['gP0'] * Pinf).sum()
        g_ls, g_var = torch.autograd.grad(surrogate, [ls, var])
        out = {'lengthscale': g_ls, 'variance': g_var}
        if isinstance(self.likelihood, Gaussian):
            out['noise'] = torch.diagonal(g['gR'], dim1=-2, dim2=-1).sum(-1)
        return lml, out

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
