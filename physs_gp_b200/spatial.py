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


def spatial_conditional_block(Kzz, Ksz, Kss, Ktt, pred_mean, pred_var, diagonal=False, jitter=None, stream=None):
    """pred_mean [T, M, 1] (or [T, M]), pred_var [T, M, M] (or [T, 1, M, M]) device or host arrays in
    time-(latent-)space format; Ktt [T], a scalar (stationary temporal kernel: its variance), or None for 1.
    Returns (mu [T, N, 1], var [T, 1, N, N]) as the reference does (:198-201), or var [T, N, 1] with `diagonal`."""
    dev = _device()
    lib = _lib.load()
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
    Wd, C0d = _to_dev(W, dev), _to_dev(C0, dev)
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
