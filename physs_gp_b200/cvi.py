"""CVI natural-gradient step and ELBO on the B200 kernels -- host-side mirror of

  stgp/approximate_posteriors/conjugate_gaussian_approximate_posterior.py:174-246  FullConjugateGaussian
  stgp/models/vgp.py:148-157,274-282              VGP.get_objective / natural_gradient_update
  stgp/computation/natural_gradients/cvi_nat_grad.py:346-410   natural_gradients(FullConjugateGaussian)
  stgp/computation/elbos/elbos.py:163-194          elbo(FullConjugateGaussian)

The sites (Y~ [T, D], V~ [T, D, D]) are literally the data and the BlockDiagonalGaussian noise of a
surrogate `SDE_GP`, exactly as in the reference; one CVI iteration is
    q_mu, q_var = surrogate.posterior_blocks()          (filter + smoother kernels)
    sites <- physs_cvi_natgrad_step_f64(sites, q_mu, q_var, data, likelihood, beta)
and the ELBO is  sum ELL(data) - sum ELL(surrogate) + lml(surrogate).
A leading batch axis B (independent blocks / series) is carried throughout.
"""
import numpy as np
import torch

from . import _lib
from . import settings
from .data import TemporalData
from .likelihood import BlockDiagonalGaussian, PrecisionBlockDiagonalGaussian
from .models import SDE_GP


# --------------------------------------------------------------------------- likelihood descriptions
class GaussianLik:
    """Gaussian likelihood on f = W u with noise covariance [P, P] (shared) -- likelihood/gaussian.py."""
    kind = _lib.LIK_GAUSS

    def __init__(self, noise):
        self.noise = np.atleast_2d(np.asarray(noise, np.float64))
        self.param = 0.0


class PoissonLik:
    """likelihood/poisson.py:9-28 (exp link)."""
    kind = _lib.LIK_POISSON_EXP

    def __init__(self, binsize=1.0):
        self.param = float(binsize)
        self.noise = None


class BernoulliLik:
    """likelihood/bernoulli.py:11-25 (probit link, +1e-5 inside the logs)."""
    kind = _lib.LIK_BERNOULLI_PROBIT

    def __init__(self):
        self.param = 0.0
        self.noise = None


_GH_CACHE = {}


def gauss_hermite(K, dev):
    key = (K, str(dev))
    if key not in _GH_CACHE:
        x, w = np.polynomial.hermite.hermgauss(K)
        _GH_CACHE[key] = (torch.as_tensor(x, device=dev), torch.as_tensor(w / np.sqrt(np.pi), device=dev))
    return _GH_CACHE[key]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _c(t):
    assert t.is_cuda and t.dtype == torch.float64
    return t.contiguous()


def _is_tm(t):
    """[B, T, ...] view of a contiguous [T, B, ...] tensor (the time-major batch layout of ops.py)."""
    return t.dim() >= 3 and not t.is_contiguous() and t.transpose(0, 1).is_contiguous()


def _block_order(tensors):
    """The site kernels treat the leading dims as one flat axis of independent blocks, so any memory order
    works as long as every per-block array uses the same one.  If all of them are time-major views, run on
    the underlying [T, B, ...] memory (no copies); otherwise make everything batch-major contiguous.
    Returns (tensors, time_major)."""
    live = [t for t in tensors if t is not None]
    if live and all(_is_tm(t) for t in live):
        return [None if t is None else t.transpose(0, 1) for t in tensors], True
    return [None if t is None else _c(t) for t in tensors], False


def _like(x, tm):
    """Uninitialised tensor with the logical shape and memory order of x (x already in kernel order)."""
    return torch.empty_like(x)


def _back(x, tm):
    return x.transpose(0, 1) if (tm and x is not None) else x


def _like_sites(x, sites):
    """x [B, T, ...] re-stored in the memory order of the site tensors (the site kernels want one order)."""
    return time_major_blocks(x) if (_is_tm(sites) and not _is_tm(x)) else x


def time_major_blocks(x):
    """Re-store a [B, T, ...] tensor time-major (logical shape unchanged)."""
    return x.transpose(0, 1).contiguous().transpose(0, 1)


def natgrad_step(Ytil, Vtil, q_mu, q_var, y, W, lik, beta, ng_jitter=None, K=20, dm=None, dS=None,
                 want_ell=False, out=None, stream=None, precision=False):
    """Raw op: all tensors are CUDA float64 with the site blocks flattened to N = prod(leading dims).
    Ytil [..., D], Vtil [..., D, D], q_mu, q_var alike; y [..., P]; W [P, D] or None.
    `precision=True`: the 'NG_Precision' parameterisation -- Vtil holds the site PRECISION in and out
    (physs_cvi_natgrad_step_prec_f64; exponential_family_transforms.py:44-53,85-95).
    Returns (Ytil_new, Vtil_new[, ell [...]])."""
    lib = _lib.load()
    if Ytil.shape[-1] > BIG_BLOCK_MIN:
        if precision:
            raise NotImplementedError("NG_Precision sites: blocks up to %d x %d" % (BIG_BLOCK_MIN, BIG_BLOCK_MIN))
        return _natgrad_step_big(Ytil, Vtil, q_mu, q_var, y, W, lik, beta, ng_jitter, dm, dS, want_ell, out, stream)
    given = dm is not None
    (Ytil, Vtil, q_mu, q_var, yv, dm, dS), tm = _block_order(
        [Ytil, Vtil, q_mu, q_var, None if given else y, dm, dS])
    D = Ytil.shape[-1]
    lead = Ytil.shape[:-1]
    N = int(np.prod(lead))
    dev = Ytil.device
    kind = _lib.LIK_GIVEN if given else lik.kind
    P = 1 if given else yv.shape[-1]
    Wv = None if W is None else _c(W)
    noise = None
    nstride = 0
    if not given and lik.kind == _lib.LIK_GAUSS:
        noise = torch.as_tensor(lik.noise, device=dev) if not isinstance(lik.noise, torch.Tensor) else lik.noise
        noise = _c(noise)
        nstride = 0 if noise.dim() == 2 else P * P
    ghx = ghw = None
    if kind in (_lib.LIK_POISSON_EXP, _lib.LIK_BERNOULLI_PROBIT):
        ghx, ghw = gauss_hermite(K, dev)
    if out is None:
        Yn, Vn = torch.empty_like(Ytil), torch.empty_like(Vtil)
    else:
        (Yn, Vn), tm_out = _block_order(list(out))
        if tm_out != tm or Yn.data_ptr() != out[0].data_ptr() or Vn.data_ptr() != out[1].data_ptr():
            raise ValueError("out buffers must use the memory order of the sites")
    ell = torch.empty(lead, dtype=torch.float64, device=dev) if want_ell else None
    ngj = settings.ng_jitter if ng_jitter is None else ng_jitter
    s = stream if stream is not None else torch.cuda.current_stream()
    step_fn = lib.physs_cvi_natgrad_step_prec_f64 if precision else lib.physs_cvi_natgrad_step_f64
    with torch.cuda.device(dev):
        st = step_fn(
            s.cuda_stream, N, D, P, kind, Ytil.data_ptr(), Vtil.data_ptr(), q_mu.data_ptr(), q_var.data_ptr(),
            _ptr(yv), _ptr(Wv), _ptr(noise), nstride, float(0.0 if given else lik.param), int(K),
            _ptr(ghx), _ptr(ghw), _ptr(dm if given else None), _ptr(dS if given else None),
            float(beta), float(ngj), Yn.data_ptr(), Vn.data_ptr(), _ptr(ell))
    _lib.check(st, "physs_cvi_natgrad_step_f64")
    Yn, Vn, ell = _back(Yn, tm), _back(Vn, tm), _back(ell, tm)
    return (Yn, Vn, ell) if want_ell else (Yn, Vn)


def expected_log_likelihood(q_mu, q_var, y, W, lik, K=20, noise=None, want_grads=False, stream=None):
    """Raw op: per-block ELL [...], optionally with dELL/dm [..., D] and dELL/dS [..., D, D].
    `noise` overrides lik.noise with a per-block tensor [..., P, P] (used for the surrogate ELL)."""
    lib = _lib.load()
    if q_mu.shape[-1] > BIG_BLOCK_MIN:
        return _expected_log_likelihood_big(q_mu, q_var, y, W, lik, noise, want_grads, stream)
    per_block_noise = noise if (noise is not None and noise.dim() > 2) else None
    (q_mu, q_var, y, per_block_noise), tm = _block_order([q_mu, q_var, y, per_block_noise])
    D, P = q_mu.shape[-1], y.shape[-1]
    lead = q_mu.shape[:-1]
    N = int(np.prod(lead))
    dev = q_mu.device
    Wv = None if W is None else _c(W)
    nz, nstride = None, 0
    if lik.kind == _lib.LIK_GAUSS:
        if per_block_noise is not None:
            nz = per_block_noise
        else:
            nz = _c(noise if noise is not None else torch.as_tensor(lik.noise, device=dev))
        nstride = 0 if nz.dim() == 2 else P * P
    ghx = ghw = None
    if lik.kind != _lib.LIK_GAUSS:
        ghx, ghw = gauss_hermite(K, dev)
    ell = torch.empty(lead, dtype=torch.float64, device=dev)
    dm = torch.empty(lead + (D,), dtype=torch.float64, device=dev) if want_grads else None
    dS = torch.empty(lead + (D, D), dtype=torch.float64, device=dev) if want_grads else None
    s = stream if stream is not None else torch.cuda.current_stream()
    with torch.cuda.device(dev):
        st = lib.physs_cvi_ell_f64(s.cuda_stream, N, D, P, lik.kind, q_mu.data_ptr(), q_var.data_ptr(),
                                   y.data_ptr(), _ptr(Wv), _ptr(nz), nstride, float(lik.param), int(K),
                                   _ptr(ghx), _ptr(ghw), ell.data_ptr(), _ptr(dm), _ptr(dS))
    _lib.check(st, "physs_cvi_ell_f64")
    ell, dm, dS = _back(ell, tm), _back(dm, tm), _back(dS, tm)
    return (ell, dm, dS) if want_grads else ell


# ------------------------------------------------------------------ large site blocks (config 2: D = Ns = 200)
BIG_BLOCK_MIN = 32            # site blocks above this size go to the one-CTA-per-block kernels (D <= 208)


def _big_ws(D, dev):
    with torch.cuda.device(dev):
        n = _lib.load().physs_cvi_big_workspace_bytes(D)
    if n <= 0:
        raise NotImplementedError("CVI site blocks: D <= 208 (D = %d)" % D)
    return torch.empty((n // 8 + 2,), dtype=torch.float64, device=dev)


def gaussian_diag_ell(q_mu, q_var, y, noise_var, want_grads=False):
    """Closed-form block ELL of a Gaussian likelihood with DIAGONAL noise on the block entries themselves (W = I):
    sum over the observed entries of log N(y_a | m_a, s2_a) - S_aa / (2 s2_a) (expected_log_likelihoods.py:90-117 with a
    diagonal R), its gradient dm = (y - m) / s2 and the DIAGONAL of dS = -1 / (2 s2), zero at missing entries.
    Elementwise torch ops on the device the marginals live on."""
    obs = ~torch.isnan(y)
    y0 = torch.where(obs, y, torch.zeros_like(y))
    s2 = noise_var.expand_as(q_mu)
    diag = torch.diagonal(q_var, dim1=-2, dim2=-1)
    r = y0 - q_mu
    ell = torch.where(obs, -0.5 * (np.log(2 * np.pi) + torch.log(s2) + (r * r + diag) / s2), torch.zeros_like(r)).sum(-1)
    if not want_grads:
        return ell
    dm = torch.where(obs, r / s2, torch.zeros_like(r))
    dS = torch.where(obs, -0.5 / s2, torch.zeros_like(r))
    return ell, dm, dS


def _diag_noise(lik, noise, D, dev):
    nz = noise if noise is not None else lik.noise
    nz = nz if isinstance(nz, torch.Tensor) else torch.as_tensor(np.asarray(nz, np.float64), device=dev)
    if nz.dim() >= 2 and nz.shape[-1] == nz.shape[-2]:
        off = nz - torch.diag_embed(torch.diagonal(nz, dim1=-2, dim2=-1))
        if bool((off != 0).any()):
            return None
        nz = torch.diagonal(nz, dim1=-2, dim2=-1)
    return nz.to(dev)


def _natgrad_step_big(Ytil, Vtil, q_mu, q_var, y, W, lik, beta, ng_jitter, dm, dS, want_ell, out, stream):
    lib = _lib.load()
    D = Ytil.shape[-1]
    lead = Ytil.shape[:-1]
    dev = Ytil.device
    ell = None
    if dm is None:
        if W is not None or getattr(lik, "kind", None) != _lib.LIK_GAUSS:
            raise NotImplementedError("site blocks with D > %d: Gaussian likelihood on the block entries, or supplied "
                                      "ELL gradients (dm, dS)" % BIG_BLOCK_MIN)
        nz = _diag_noise(lik, None, D, dev)
        if nz is None:
            raise NotImplementedError("site blocks with D > %d: diagonal likelihood noise" % BIG_BLOCK_MIN)
        ell, dm, dS = gaussian_diag_ell(q_mu, q_var, y, nz, want_grads=True)
    Ytil, Vtil, q_mu, dm, dS = (_c(x) for x in (Ytil, Vtil, q_mu, dm, dS))
    N = int(np.prod(lead))
    diag = dS.dim() == dm.dim()
    if out is None:
        Yn, Vn = torch.empty_like(Ytil), torch.empty_like(Vtil)
    else:
        Yn, Vn = torch.empty_like(Ytil), torch.empty_like(Vtil)       # the kernel reads the old sites while writing
    ws = _big_ws(D, dev)
    ngj = settings.ng_jitter if ng_jitter is None else ng_jitter
    s = stream if stream is not None else torch.cuda.current_stream()
    with torch.cuda.device(dev):
        st = lib.physs_cvi_natgrad_big_f64(s.cuda_stream, N, D, Ytil.data_ptr(), Vtil.data_ptr(), q_mu.data_ptr(),
                                           dm.data_ptr(), dS.data_ptr(), 1 if diag else 0, float(beta), float(ngj),
                                           ws.data_ptr(), ws.numel() * 8, Yn.data_ptr(), Vn.data_ptr())
    _lib.check(st, "physs_cvi_natgrad_big_f64")
    if out is not None:
        out[0].copy_(Yn); out[1].copy_(Vn)
        Yn, Vn = out
    return (Yn, Vn, ell) if want_ell else (Yn, Vn)


def _expected_log_likelihood_big(q_mu, q_var, y, W, lik, noise, want_grads, stream):
    """D > BIG_BLOCK_MIN: diagonal Gaussian likelihood in closed form (torch), or -- `noise` a per-block full
    covariance [..., D, D], the surrogate ELL -- physs_cvi_ell_sur_big_f64."""
    lib = _lib.load()
    D = q_mu.shape[-1]
    dev = q_mu.device
    if getattr(lik, "kind", None) != _lib.LIK_GAUSS or W is not None:
        raise NotImplementedError("site blocks with D > %d: Gaussian likelihoods on the block entries" % BIG_BLOCK_MIN)
    nz = _diag_noise(lik, noise, D, dev)
    if nz is not None:
        return gaussian_diag_ell(q_mu, q_var, y, nz, want_grads=want_grads)
    if want_grads:
        raise NotImplementedError("gradients of a full-covariance block ELL with D > %d" % BIG_BLOCK_MIN)
    lead = q_mu.shape[:-1]
    N = int(np.prod(lead))
    Yt, Vt, qm, qS = (_c(x) for x in (y, noise, q_mu, q_var))
    ell = torch.empty(lead, dtype=torch.float64, device=dev)
    ws = _big_ws(D, dev)
    s = stream if stream is not None else torch.cuda.current_stream()
    with torch.cuda.device(dev):
        st = lib.physs_cvi_ell_sur_big_f64(s.cuda_stream, N, D, Yt.data_ptr(), Vt.data_ptr(), qm.data_ptr(), qS.data_ptr(),
                                           ws.data_ptr(), ws.numel() * 8, ell.data_ptr())
    _lib.check(st, "physs_cvi_ell_sur_big_f64")
    return ell


class DampedPendulumLik:
    """Likelihood of the PHYSS-GP damped-oscillator model: MultiOutput([observe x, DampedPendulum1D residual])
    (zoo/sde_diff.py:757-763, transforms/pdes.py:530-597) with Gaussian noise `var_obs` on x and `var_col` on
    the collocation residual.  Data rows are (observation of x, collocation target = 0), NaN = absent.
    state_index: positions of (x, x_t, x_tt) inside the site block."""
    kind = "pendulum"

    def __init__(self, g, l, b, var_obs, var_col, state_index=(0, 1, 2)):
        self.g_over_l, self.b = float(g) / float(l), float(b)
        self.var_obs, self.var_col = float(var_obs), float(var_col)
        self.state_index = tuple(int(i) for i in state_index)


def pendulum_expected_log_likelihood(q_mu, q_var, y, lik, gauss_newton=False, want_grads=False, stream=None):
    """Raw op: closed-form collocation ELL [...], optionally dELL/dm [..., D] and the site curvature
    [..., D, D] (exact dELL/dS, or its Gauss-Newton delta-u replacement).  y [..., 2]."""
    lib = _lib.load()
    (q_mu, q_var, y), tm = _block_order([q_mu, q_var, y])
    D = q_mu.shape[-1]
    lead = q_mu.shape[:-1]
    N = int(np.prod(lead))
    dev = q_mu.device
    ell = torch.empty(lead, dtype=torch.float64, device=dev)
    dm = torch.empty(lead + (D,), dtype=torch.float64, device=dev) if want_grads else None
    dS = torch.empty(lead + (D, D), dtype=torch.float64, device=dev) if want_grads else None
    s = stream if stream is not None else torch.cuda.current_stream()
    i0, i1, i2 = lik.state_index
    with torch.cuda.device(dev):
        st = lib.physs_cvi_ell_pendulum_f64(s.cuda_stream, N, D, i0, i1, i2, q_mu.data_ptr(), q_var.data_ptr(),
                                            y.data_ptr(), lik.g_over_l, lik.b, lik.var_obs, lik.var_col,
                                            1 if gauss_newton else 0, ell.data_ptr(), _ptr(dm), _ptr(dS))
    _lib.check(st, "physs_cvi_ell_pendulum_f64")
    ell, dm, dS = _back(ell, tm), _back(dm, tm), _back(dS, tm)
    return (ell, dm, dS) if want_grads else ell


def gauss_newton_curvature(J, var, y=None, stream=None):
    """Raw op: dS [..., D, D] = -1/2 sum_p mask_p J_p^T J_p / var_p from Jacobians J [..., P, D] of the prior
    transform (computed by the caller, e.g. jax.jacfwd), likelihood variances var [P] or [..., P], and optional
    data y [..., P] whose NaNs mask outputs (cvi_hessian_approximations.py:380-431)."""
    lib = _lib.load()
    J = _c(J)
    P, D = J.shape[-2], J.shape[-1]
    lead = J.shape[:-2]
    N = int(np.prod(lead))
    var = _c(torch.as_tensor(var, dtype=torch.float64, device=J.device))
    vstride = 0 if var.dim() == 1 else P
    yv = None if y is None else _c(y)
    dS = torch.empty(lead + (D, D), dtype=torch.float64, device=J.device)
    s = stream if stream is not None else torch.cuda.current_stream()
    with torch.cuda.device(J.device):
        st = lib.physs_cvi_gauss_newton_f64(s.cuda_stream, N, D, P, J.data_ptr(), var.data_ptr(), vstride, _ptr(yv),
                                            dS.data_ptr())
    _lib.check(st, "physs_cvi_gauss_newton_f64")
    return dS


# --------------------------------------------------------------------------- reference-shaped objects
class FullConjugateGaussian:
    """q(u) prop. to N(Y~ | u, V~) p(u): sites stored as the surrogate SDE_GP's data and noise
    (conjugate_gaussian_approximate_posterior.py:174-246).  Reference initialisation: Y~ = 1e-5,
    V~ = I (:209-218)."""

    def __init__(self, X_time, surrogate_prior, block_size, B=1, Y_tilde=None, V_tilde=None, device=None,
                 filter_type='b200', parameterisation='NG_Moment'):
        if parameterisation not in ('NG_Moment', 'NG_Precision'):
            raise ValueError("parameterisation: 'NG_Moment' or 'NG_Precision' (natural_gradients/parameterisations.py)")
        # 'NG_Precision': V_tilde holds the site PRECISION (the surrogate's PrecisionBlockDiagonalGaussian); the
        # reference's initialisation theta_2 = I is the same matrix either way
        self.parameterisation = parameterisation
        dev = device or torch.device("cuda", torch.cuda.current_device())
        T = len(X_time)
        D = block_size
        self.block_size = D
        self.X_time = X_time
        self.Y_tilde = (torch.full((B, T, D), 1e-5, dtype=torch.float64, device=dev) if Y_tilde is None
                        else torch.as_tensor(Y_tilde, dtype=torch.float64).to(dev).reshape(B, T, D).clone())
        self.V_tilde = (torch.eye(D, dtype=torch.float64, device=dev).expand(B, T, D, D).contiguous()
                        if V_tilde is None
                        else torch.as_tensor(V_tilde, dtype=torch.float64).to(dev).reshape(B, T, D, D).clone())
        if settings.time_major and B >= settings.time_major_min_batch:
            # sites live in the batch layout the filter / smoother kernels produce (no re-ordering copies)
            self.Y_tilde, self.V_tilde = time_major_blocks(self.Y_tilde), time_major_blocks(self.V_tilde)
        self.prior = surrogate_prior
        self.filter_type = filter_type        # threaded to the surrogate SDE_GP (zoo/sde_diff.py:690,712,736,752)

    @property
    def surrogate(self):
        data = TemporalData(self.X_time, self.Y_tilde[..., None])          # [B, T, P=D, Ns=1]
        lik = (PrecisionBlockDiagonalGaussian(self.V_tilde) if self.parameterisation == 'NG_Precision'
               else BlockDiagonalGaussian(self.V_tilde))
        return SDE_GP(data, self.prior, lik, filter_type=self.filter_type)


class VGP:
    """models/vgp.py: variational GP with a CVI approximate posterior.
    data Y [B, T, P] (NaN = missing); W [P, D] maps a site block to the likelihood inputs
    (None = identity)."""

    def __init__(self, Y, likelihood, approximate_posterior, W=None, ell_quad_points=20):
        q = approximate_posterior
        dev = q.Y_tilde.device
        self.q = q
        self.set_data(Y)
        self.lik = likelihood
        self.W = None if W is None else torch.as_tensor(W, dtype=torch.float64).to(dev)
        self.K = ell_quad_points

    def set_data(self, Y):
        """Data [B, T, P] (host or device); stored in the memory order of the sites.  A second call with the same
        shape overwrites the existing device buffer in place (a compiled step keeps reading it)."""
        dev = self.q.Y_tilde.device
        Y = torch.as_tensor(Y, dtype=torch.float64).to(dev, non_blocking=True)
        if Y.dim() == 2:
            Y = Y[None]
        cur = getattr(self, "Y", None)
        if cur is not None and cur.shape == Y.shape:
            cur.copy_(Y)                       # strided destination when the sites are time-major
            return
        self.Y = time_major_blocks(Y) if _is_tm(self.q.Y_tilde) else Y

    def stage_data(self, Y):
        """Start uploading the NEXT data set [B, T, P] (pinned host memory) on a side stream into a staging buffer while
        the current iteration computes; `commit_data()` then swaps it in.  The pair does what `set_data` does, with the
        host -> device copy overlapped with the previous step (the upload of 80 MB takes as long as the step itself)."""
        dev = self.q.Y_tilde.device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = torch.empty(self.Y.shape, dtype=torch.float64, device=dev)   # dense [B, T, P]
            self._staged, self._consumed = torch.cuda.Event(), None
        Y = torch.as_tensor(Y, dtype=torch.float64)
        if Y.dim() == 2:
            Y = Y[None]
        with torch.cuda.stream(self._copy_stream):
            if self._consumed is not None:
                self._copy_stream.wait_event(self._consumed)      # the previous commit has read the staging buffer
            self._staging.copy_(Y, non_blocking=True)
            self._staged.record(self._copy_stream)

    def commit_data(self):
        """Make the staged data set current (device -> device, into the buffer a compiled step reads)."""
        cur = torch.cuda.current_stream(self.q.Y_tilde.device)
        cur.wait_event(self._staged)
        self.Y.copy_(self._staging)                # strided destination when the sites are time-major
        self._consumed = torch.cuda.Event()
        self._consumed.record(cur)

    def compile_step(self, lr, enforce_psd_type=None, reuse_posterior=False):
        """Capture natural_gradient_update(lr) + elbo() -- one CVI iteration, vgp.py:274-282,148-157 -- into ONE CUDA
        graph.  The iteration is ~85 short launches (chunk summaries, scans, replays, site kernels, reductions); as
        a graph the host enqueues a single node list and the launch gaps disappear.  The sites, the data buffer
        (set_data) and the returned ELBO tensor keep their addresses; step() replays the graph.  Parallel-in-time
        filters cannot fall back inside a graph: their device flags are OR-ed into self.step_status (0 = converged),
        which the caller may read whenever a host sync is acceptable."""
        from . import filters as _filters
        q = self.q
        Yt0, Vt0 = q.Y_tilde.clone(), q.V_tilde.clone()
        side = torch.cuda.Stream(device=q.Y_tilde.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the capture: caches, workspaces, kernel attributes
            for _ in range(2):
                self.natural_gradient_update(lr, enforce_psd_type=enforce_psd_type)
                self.elbo()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        q.Y_tilde.copy_(Yt0); q.V_tilde.copy_(Vt0)
        self._post = None
        if reuse_posterior:
            # the posterior under the CURRENT sites, in buffers that keep their addresses: every replay reads them in
            # its natural-gradient half and refreshes them from the posterior its ELBO half computes
            self.reuse_posterior = True
            self._posterior(want_lml=True)
            self._post_static = tuple(x.clone() for x in self._post)
            self._post = self._post_static
            torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        del _filters.captured_status[:]
        with torch.cuda.graph(graph):
            self.natural_gradient_update(lr, enforce_psd_type=enforce_psd_type)
            elbo = self.elbo()
            if reuse_posterior:
                for dst, src in zip(self._post_static, self._post):
                    dst.copy_(src)
                self._post = self._post_static
            flags = list(_filters.captured_status)
            status = torch.stack([f.reshape(-1)[0] for f in flags]).max() if flags else None
        del _filters.captured_status[:]
        self._graph, self._graph_elbo, self.step_status = graph, elbo, status
        return self

    def step(self):
        """Replay the compiled iteration (compile_step); returns the ELBO tensor [B] (same storage every call)."""
        self._graph.replay()
        return self._graph_elbo

    reuse_posterior = False     # opt-in: see _posterior

    def _posterior(self, want_lml=False):
        """(lml, q_mu [B,T,D], q_var [B,T,D,D]) of the surrogate under the current sites.  The reference runs the
        filter + smoother once in `natural_gradients` and once more in `elbo` (vgp.py:274-282, 148-157); in a training
        loop the pass of iteration i's ELBO and the one of iteration i + 1's natural-gradient step see the SAME sites.
        With `reuse_posterior = True` the result is kept until the sites change (natural_gradient_update) and served
        from there: one posterior pass per iteration instead of two, same numbers.  The surrogate's prior must not
        change in between (call `invalidate()` after a hyper-parameter step)."""
        if self.reuse_posterior and getattr(self, "_post", None) is not None:
            return self._post
        lml, q_mu, q_var = self.q.surrogate.posterior_blocks(return_lml=True)
        out = (lml, q_mu[..., 0], q_var[..., 0, :, :])
        if self.reuse_posterior:
            self._post = out
        return out

    def invalidate(self):
        self._post = None

    def natural_gradient_update(self, lr, enforce_psd_type=None, prediction_samples=None):
        """vgp.py:274-282 -> cvi_nat_grad.py:508-515,346-410 -> cvi_parameterisations.py:63-93."""
        pde = getattr(self.lik, "kind", None) == "pendulum"
        if enforce_psd_type not in (None, 'laplace_gauss_newton_delta_u') or (enforce_psd_type and not pde):
            raise NotImplementedError("enforce_psd_type: None, or 'laplace_gauss_newton_delta_u' with a PDE "
                                      "collocation likelihood, are implemented on the b200 path")
        q = self.q
        prec = getattr(q, "parameterisation", 'NG_Moment') == 'NG_Precision'   # cvi_parameterisations.py:95-113
        _, q_mu, q_var = self._posterior()                         # [B,T,D], [B,T,D,D]
        self._post = None                                          # the sites change below
        if pde:
            # cvi_nat_grad.py:381-387: dELL/dm from the ELL, dELL/dS replaced by the Gauss-Newton curvature
            _, dm, dS = pendulum_expected_log_likelihood(q_mu, q_var, self.Y, self.lik,
                                                         gauss_newton=enforce_psd_type is not None, want_grads=True)
            natgrad_step(q.Y_tilde, q.V_tilde, q_mu, q_var, None, None, None, lr, dm=dm, dS=dS,
                         out=(q.Y_tilde, q.V_tilde), precision=prec)
            return
        natgrad_step(q.Y_tilde, q.V_tilde, q_mu, q_var, self.Y, self.W, self.lik, lr, K=self.K,
                     out=(q.Y_tilde, q.V_tilde), precision=prec)

    def elbo(self):
        """elbos.py:163-194; returns one ELBO per batch member [B]."""
        q = self.q
        lml, q_mu, q_var = self._posterior(want_lml=True)
        if getattr(self.lik, "kind", None) == "pendulum":
            ell = pendulum_expected_log_likelihood(q_mu, q_var, self.Y, self.lik)
        else:
            ell = expected_log_likelihood(q_mu, q_var, self.Y, self.W, self.lik, K=self.K)
        sur = GaussianLik(np.eye(q.block_size))
        noise = q.V_tilde
        if getattr(q, "parameterisation", 'NG_Moment') == 'NG_Precision':
            noise = _like_sites(PrecisionBlockDiagonalGaussian(q.V_tilde).variance, q.V_tilde)   # mat_inv(precision)
        ell_s = expected_log_likelihood(q_mu, q_var, q.Y_tilde, None, sur, noise=noise)
        return sum_steps(ell, ell_s) + lml

    def get_objective(self):
        return -self.elbo()


def sum_steps(x, sub=None, stream=None):
    """out[b] = sum_k (x[b, k] - sub[b, k]) for [B, T] device tensors in either step layout (`physs_sum_steps_f64`): the
    ELL sums of the ELBO in one deterministic two-stage pass."""
    if x.dim() != 2 or (sub is not None and (sub.shape != x.shape or sub.stride() != x.stride())):
        return x.sum(dim=-1) - (0 if sub is None else sub.sum(dim=-1))
    B, T = x.shape
    sb, st_ = x.stride()
    if not ((sb == T and st_ == 1) or (sb == 1 and st_ == B)):
        return x.sum(dim=-1) - (0 if sub is None else sub.sum(dim=-1))
    out = torch.empty((B,), dtype=torch.float64, device=x.device)
    scratch = torch.empty((B * ((T + 511) // 512),), dtype=torch.float64, device=x.device)
    s = stream if stream is not None else torch.cuda.current_stream()
    with torch.cuda.device(x.device):
        rc = _lib.load().physs_sum_steps_f64(s.cuda_stream, B, T, sb, st_, x.data_ptr(), _ptr(sub), scratch.data_ptr(),
                                             out.data_ptr())
    _lib.check(rc, "physs_sum_steps_f64")
    return out


class MeanFieldConjugateGaussian:
    """q(u) = prod_q q_q(u_q): one conjugate-Gaussian factor -- i.e. one surrogate SDE_GP -- per latent
    (approximate_posteriors/conjugate_gaussian_approximate_posterior.py:127-172).  The reference runs the
    Q surrogates under batch_or_loop (cvi_nat_grad_utils.py:111-119); here they are Q FullConjugateGaussian
    objects whose filter / smoother calls are independent batched launches."""

    def __init__(self, approx_posteriors):
        self.approx_posteriors = list(approx_posteriors)
        self.block_sizes = [q.block_size for q in self.approx_posteriors]
        self.block_size = sum(self.block_sizes)


class MeanFieldVGP:
    """VGP with a mean-field CVI posterior (models/vgp.py with MeanFieldConjugateGaussian):
    natural_gradients mean-field variant cvi_nat_grad.py:89-145 and elbo elbos.py:136-160.

    The likelihood sees the stacked latent block u = (u_1, ..., u_Q) through W [P, sum D_q]; under the
    mean-field posterior its covariance is block diagonal, the ELL gradients are evaluated on the stacked
    marginal by the same kernel as the full posterior, and each factor is updated with its own diagonal
    block of dELL/dS (cvi_nat_grad.py:103-139)."""

    def __init__(self, Y, likelihood, approximate_posterior, W=None, ell_quad_points=20):
        self.q = approximate_posterior
        q0 = self.q.approx_posteriors[0]
        dev = q0.Y_tilde.device
        Y = torch.as_tensor(Y, dtype=torch.float64).to(dev)
        self.Y = Y[None] if Y.dim() == 2 else Y
        self.lik = likelihood
        self.W = None if W is None else torch.as_tensor(W, dtype=torch.float64).to(dev)
        self.K = ell_quad_points

    def _stacked_marginal(self, want_lml=False):
        mus, covs, lmls = [], [], []
        for q in self.q.approx_posteriors:
            lml, mu, var = q.surrogate.posterior_blocks(return_lml=True)
            mus.append(mu[..., 0].contiguous())
            covs.append(var[..., 0, :, :].contiguous())
            lmls.append(lml)
        m = torch.cat(mus, dim=-1)
        D = m.shape[-1]
        S = torch.zeros(m.shape + (D,), dtype=torch.float64, device=m.device)
        o = 0
        for c in covs:
            k = c.shape[-1]
            S[..., o:o + k, o:o + k] = c
            o += k
        return (m, S, mus, covs, lmls) if want_lml else (m, S, mus, covs)

    def natural_gradient_update(self, lr, enforce_psd_type=None, prediction_samples=None):
        if enforce_psd_type is not None:
            raise NotImplementedError("enforce_psd_type is not implemented for the mean-field posterior")
        m, S, mus, covs = self._stacked_marginal()
        _, dm, dS = expected_log_likelihood(m, S, self.Y, self.W, self.lik, K=self.K, want_grads=True)
        o = 0
        for q, mu_q, cov_q in zip(self.q.approx_posteriors, mus, covs):
            k = q.block_size
            dm_q = dm[..., o:o + k].contiguous()
            dS_q = dS[..., o:o + k, o:o + k].contiguous()
            Yt, Vt = q.Y_tilde.contiguous(), q.V_tilde.contiguous()
            # each factor keeps its own parameterisation (cvi_parameterisations.py:14-61)
            Yn, Vn = natgrad_step(Yt, Vt, mu_q, cov_q, None, None, None, lr, dm=dm_q, dS=dS_q,
                                  precision=getattr(q, "parameterisation", 'NG_Moment') == 'NG_Precision')
            q.Y_tilde.copy_(Yn)
            q.V_tilde.copy_(Vn)
            o += k

    def elbo(self):
        m, S, mus, covs, lmls = self._stacked_marginal(want_lml=True)
        ell = expected_log_likelihood(m, S, self.Y, self.W, self.lik, K=self.K).sum(dim=-1)
        out = ell
        for q, mu_q, cov_q, lml in zip(self.q.approx_posteriors, mus, covs, lmls):
            sur = GaussianLik(np.eye(q.block_size))
            noise = q.V_tilde.contiguous()
            if getattr(q, "parameterisation", 'NG_Moment') == 'NG_Precision':
                noise = PrecisionBlockDiagonalGaussian(noise).variance
            ell_s = expected_log_likelihood(mu_q, cov_q, q.Y_tilde.contiguous(), None, sur, noise=noise)
            out = out - ell_s.sum(dim=-1) + lml           # - KL_q = -(ELL_sur,q - lml_q)   (elbos.py:74-90)
        return out

    def get_objective(self):
        return -self.elbo()
