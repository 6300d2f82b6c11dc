"""Spatial conditional after the smoother (second half) -- mirror of
  stgp/computation/spatial_conditionals.py:30-207  spatial_conditional_block (the f_only branch) and
  stgp/computation/spatial_conditionals.py:209-229 spatial_conditional
for separable spatio-temporal priors: the posterior (m_t, P_t) of f at the M inducing / training spatial points of every
time step is carried to N new spatial points,
    mu_t = Ksz Kzz^-1 m_t,     var_t = Ktt_t (Kss - Ksz Kzz^-1 Kzs) + Ksz Kzz^-1 (P_t + jitter I) Kzz^-1 Kzs.
The time-invariant pieces (one M x M Cholesky, two triangular solves: marginals.py:101-102) are evaluated once on the host
in numpy like every other T-independent quantity of the prior; the per-step work runs in
`physs_spatial_conditional_f64` (csrc/physs_kron.cu), one persistent CTA per SM.  The reference evaluates the Gram
matrices with `_batched_st_kernel` from its spatial kernel objects; spatial kernels are outside the path, so the
Gram matrices themselves are the arguments here (same names as the reference's locals)."""
import numpy as np
import scipy.linalg as sla
import torch

from . import _lib
from . import settings
from .filters import _device, _to_dev


def conditional_weights(Kzz, Ksz, Kss, jitter=None):
    """(W, C0) = (Ksz Kzz^-1, Kss - Ksz Kzz^-1 Kzs) through the Cholesky factor of Kzz + jitter I, in the reference's
    order of operations (spatial_conditionals.py:146 + marginals.py:101-102, 107)."""
    jit = settings.jitter if jitter is None else jitter
    Kzz, Ksz, Kss = (np.asarray(x, np.float64) for x in (Kzz, Ksz, Kss))
    L = np.linalg.cholesky(Kzz + jit * np.eye(Kzz.shape[0]))
    A = sla.solve_triangular(L, Ksz.T, lower=True)
    A1 = sla.solve_triangular(L.T, A, lower=False)
    return np.ascontiguousarray(A1.T), Kss - A.T @ A


def _apply(Wd, C0d, ktt, m, P, jit, diagonal, stream=None):
    """physs_spatial_conditional_f64 on device tensors: Wd [N, M], C0d [N, N], ktt [T] or None, m [T, M], P [T, M, M]."""
    lib = _lib.load()
    dev = Wd.device
    N, M = Wd.shape
    T = m.shape[0]
    mu = torch.empty((T, N, 1), dtype=torch.float64, device=dev)
    var = torch.empty((T, N, 1) if diagonal else (T, 1, N, N), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        nws = lib.physs_spatial_conditional_ws_bytes(M, N)
    if nws <= 0:
        raise _lib.PhyssError("physs_spatial_conditional_ws_bytes: no workspace size for M = %d, N = %d" % (M, N))
    ws = torch.empty((nws // 8 + 2,), dtype=torch.float64, device=dev)
    s = stream if stream is not None else torch.cuda.current_stream()
    with torch.cuda.device(dev):
        st = lib.physs_spatial_conditional_f64(s.cuda_stream, T, M, N, Wd.data_ptr(), C0d.data_ptr(),
                                               None if ktt is None else ktt.data_ptr(), m.data_ptr(), P.data_ptr(),
                                               float(jit), 1 if diagonal else 0, ws.data_ptr(), ws.numel() * 8,
                                               mu.data_ptr(), var.data_ptr())
    _lib.check(st, "physs_spatial_conditional_f64")
    return mu, var


def spatial_conditional_block(Kzz, Ksz, Kss, Ktt, pred_mean, pred_var, diagonal=False, jitter=None, stream=None):
    """pred_mean [T, M, 1] (or [T, M]), pred_var [T, M, M] (or [T, 1, M, M]) device or host arrays in
    time-(latent-)space format; Ktt [T], a scalar (stationary temporal kernel: its variance), or None for 1.
    Returns (mu [T, N, 1], var [T, 1, N, N]) as the reference does (:198-201), or var [T, N, 1] with `diagonal`."""
    dev = _device()
    jit = settings.jitter if jitter is None else jitter
    W, C0 = conditional_weights(Kzz, Ksz, Kss, jit)
    N, M = W.shape
    m = _to_dev(pred_mean, dev).reshape(-1, M).contiguous()
    T = m.shape[0]
    P = _to_dev(pred_var, dev).reshape(T, M, M).contiguous()
    ktt = None
    if Ktt is not None:
        k = np.asarray(Ktt.detach().cpu().numpy() if isinstance(Ktt, torch.Tensor) else Ktt, np.float64).reshape(-1)
        if k.size == 1:
            C0 = C0 * k[0]
        else:
            if k.size != T:
                raise ValueError("Ktt must have one entry per time step")
            ktt = _to_dev(k, dev)
    return _apply(_to_dev(W, dev), _to_dev(C0, dev), ktt, m, P, jit, diagonal, stream)


def spatial_conditional(model, XS_space, spatial_kernel, diagonal=True):
    """`spatial_conditional(data_xs, data_x, pred_mean, pred_var, gp, diagonal)` of the reference for a model whose prior
    is a `SpatioTemporalSeperableKernel`: smooth at the training times (f only), then carry every step's posterior from
    the model's spatial points `model.data.X_space` to `XS_space` [N, D].  `spatial_kernel(X1, X2)` evaluates the spatial
    Gram matrix (the reference's `_batched_st_kernel(.., 'spatial')`)."""
    X = np.asarray(model.data.X_space, np.float64)
    XS = np.asarray(XS_space, np.float64)
    Kzz, Ksz, Kss = spatial_kernel(X, X), spatial_kernel(XS, X), spatial_kernel(XS, XS)
    mu_t, var_t = model.filter_and_smooth(full_state=False)
    kt = _temporal_variance(model.prior)
    return spatial_conditional_block(Kzz, Ksz, Kss, kt, mu_t, var_t, diagonal=diagonal)


def _temporal_variance(prior):
    k = prior
    for attr in ("gp", "parent"):
        while hasattr(k, attr):
            k = getattr(k, attr)
            if isinstance(k, (list, tuple)):
                k = k[0]
    k = getattr(k, "kernel", k)
    k = getattr(k, "k1", k)
    return float(k.variance)


# ------------------------------------------------------------------------- SpatialSparsity CVI (Gaussian likelihood)
class SpatialSparsityVGP:
    """CVI with `SpatialSparsity` (sparsity/sparsity.py; `natural_gradients(VGP, FullConjugateGaussian, SpatialSparsity)`,
    natural_gradients/cvi_nat_grad.py:346-410): the sites -- one block per time step -- live at M inducing spatial points
    Z while the data sit at N spatial points X.  The reference evaluates the ELL on the marginals of f at X, obtained
    from q at Z with the spatial conditional (elbos/marginals/dispatched_marginal_predictors.py -> spatial_conditional),
    and takes `jax.grad` of it with respect to the block moments (:381-383).  Here, for a Gaussian likelihood with
    diagonal noise:
        m_x = W m_z,   v_x = diag(Ktt C0 + W (S_z + jitter I) W^T)                 (physs_spatial_conditional_f64)
        ELL_t = sum_i log N(y_i | m_x,i, s2) - v_x,i / (2 s2)                        (observed entries)
        dELL/dm_z = W^T (y - m_x) / s2,   dELL/dS_z = W^T diag(-1 / (2 s2)) W      (the same kernel with the roles of
                                                                                    Z and X exchanged)
    followed by the ordinary block update (physs_cvi_natgrad_big_f64 or physs_cvi_natgrad_step_f64, gradients supplied).
    With `likelihood` = cvi.PoissonLik / cvi.BernoulliLik (independent non-Gaussian observations at X) the per-point
    terms E[l(f_i)], E[l'(f_i)], 1/2 E[l''(f_i)] under f_i ~ N(m_x,i, v_x,i) come from the Gauss-Hermite site kernel
    (physs_cvi_ell_f64 on N scalar blocks per step) and are pulled back the same way:
        dELL/dm_z = W^T E[l'],   dELL/dS_z = W^T diag(1/2 E[l'']) W.
    Y [T, N] (NaN = missing); approximate_posterior: cvi.FullConjugateGaussian with block_size M and B = 1."""

    def __init__(self, Y, noise_var, approximate_posterior, Kzz, Kxz, Kxx, Ktt=1.0, jitter=None, likelihood=None,
                 ell_quad_points=20):
        q = approximate_posterior
        dev = q.Y_tilde.device
        self.q = q
        self.jitter = settings.jitter if jitter is None else jitter
        W, C0 = conditional_weights(Kzz, Kxz, Kxx, self.jitter)
        self.N, self.M = W.shape
        if q.block_size != self.M or q.Y_tilde.shape[0] != 1:
            raise ValueError("SpatialSparsityVGP: one site block of size M = %d per time step, B = 1" % self.M)
        self.W = torch.as_tensor(W, device=dev)
        self.Wt = torch.as_tensor(np.ascontiguousarray(W.T), device=dev)
        self.C0 = torch.as_tensor(C0 * float(Ktt), device=dev)
        self.zero_MM = torch.zeros((self.M, self.M), dtype=torch.float64, device=dev)
        self.Y = torch.as_tensor(np.asarray(Y, np.float64), device=dev).reshape(-1, self.N)
        self.noise_var = None if noise_var is None else float(noise_var)
        self.lik = likelihood                  # None: Gaussian with variance noise_var
        self.K = ell_quad_points
        if (self.lik is None) == (self.noise_var is None):
            raise ValueError("SpatialSparsityVGP: give noise_var (Gaussian) or likelihood (Poisson / Bernoulli), not both")

    def _ell(self, q_mu, q_var, want_grads):
        """q_mu [T, M], q_var [T, M, M] -> ELL per step [T] (and the gradients w.r.t. the block moments)."""
        m_x, v_x = _apply(self.W, self.C0, None, q_mu, q_var, self.jitter, True)
        m_x, v_x = m_x[..., 0], v_x[..., 0]
        if self.lik is not None:
            from . import cvi
            out = cvi.expected_log_likelihood(m_x[..., None].contiguous(), v_x[..., None, None].contiguous(),
                                              self.Y[..., None].contiguous(), None, self.lik, K=self.K,
                                              want_grads=want_grads)
            if not want_grads:
                return out.sum(-1)
            ell, dm_x, dS_x = out
            dm_z, dS_z = _apply(self.Wt, self.zero_MM, None, dm_x[..., 0].contiguous(),
                                torch.diag_embed(dS_x[..., 0, 0]).contiguous(), 0.0, False)
            return ell.sum(-1), dm_z[..., 0], dS_z[:, 0]
        s2 = self.noise_var
        obs = ~torch.isnan(self.Y)
        r = torch.where(obs, self.Y, torch.zeros_like(self.Y)) - m_x
        ell = torch.where(obs, -0.5 * (np.log(2 * np.pi * s2) + (r * r + v_x) / s2), torch.zeros_like(r)).sum(-1)
        if not want_grads:
            return ell
        dm_x = torch.where(obs, r / s2, torch.zeros_like(r)).contiguous()
        dS_x = torch.diag_embed(torch.where(obs, torch.full_like(r, -0.5 / s2), torch.zeros_like(r))).contiguous()
        dm_z, dS_z = _apply(self.Wt, self.zero_MM, None, dm_x, dS_x, 0.0, False)
        return ell, dm_z[..., 0], dS_z[:, 0]

    def _posterior(self, want_lml=False):
        out = self.q.surrogate.posterior_blocks(return_lml=want_lml)
        lml = out[0] if want_lml else None
        q_mu, q_var = out[-2][0, :, :, 0].contiguous(), out[-1][0, :, 0].contiguous()
        return lml, q_mu, q_var

    def natural_gradient_update(self, lr, enforce_psd_type=None, prediction_samples=None):
        if enforce_psd_type is not None:
            raise NotImplementedError("enforce_psd_type: the Gaussian ELL curvature is already negative semi-definite; "
                                      "log-concave Poisson / Bernoulli sites give a negative semi-definite W^T diag W too")
        from . import cvi
        q = self.q
        _, q_mu, q_var = self._posterior()
        _, dm, dS = self._ell(q_mu, q_var, True)
        cvi.natgrad_step(q.Y_tilde, q.V_tilde, q_mu[None], q_var[None], None, None, None, lr,
                         dm=dm[None].contiguous(), dS=dS[None].contiguous(), out=(q.Y_tilde, q.V_tilde))

    def elbo(self):
        from . import cvi
        q = self.q
        lml, q_mu, q_var = self._posterior(want_lml=True)
        ell = self._ell(q_mu, q_var, False)
        sur = cvi.GaussianLik(np.eye(q.block_size))
        ell_s = cvi.expected_log_likelihood(q_mu[None], q_var[None], q.Y_tilde, None, sur, noise=q.V_tilde)
        return ell.sum(dim=-1) - ell_s.sum(dim=-1) + lml

    def get_objective(self):
        return -self.elbo()
