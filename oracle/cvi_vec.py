"""Oracle: one whole CVI iteration (natural-gradient site update + ELBO) for SCALAR sites (D = 1), vectorised
over blocks and time steps -- the CPU restatement timed beside the B200 CVI step (bench.py `cpu_baseline` of
the `cvi` workload, BASELINE config 4).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Composition (paths relative to /root/reference/src/lib/stgp/):
  models/vgp.py:274-282            natural_gradient_update -> computation/natural_gradients/cvi_nat_grad.py:346-410
  models/vgp.py:148-157            get_objective           -> computation/elbos/elbos.py:163-194
The posterior of the surrogate model (sde_gp.py:255-277) comes from the C port of the sequential filter /
smoother (oracle/ssm_oracle.c); the per-site algebra is exactly oracle/cvi.py's (theta <-> lambda with
ng_jitter, cvi_block_update, Gauss-Hermite ELL, surrogate ELL) written for D = 1 so that numpy can run it over
[B, T] at once; tests/test_oracle_cvi.py holds it to the per-block loops of oracle/cvi.py.
"""
import numpy as np
from scipy.special import gammaln, ndtr

from . import c_oracle
from . import linalg as la

LOG2PI = np.log(2.0 * np.pi)


def gh_ell_and_grads(y, m, v, kind, K=20, binsize=1.0):
    """oracle/cvi.py:gh_ell_and_grads over arrays y, m, v [...]; NaN y -> 0."""
    x, w = np.polynomial.hermite.hermgauss(K)
    w = w / np.sqrt(np.pi)
    obs = ~np.isnan(y)
    ys = np.where(obs, y, 0.0)[..., None]
    f = m[..., None] + np.sqrt(2.0 * v)[..., None] * x
    if kind == "poisson":
        lam = np.exp(f) * binsize
        l = ys * np.log(lam) - lam - gammaln(ys + 1.0)
        d1, d2 = ys - lam, -lam
    elif kind == "bernoulli":
        p = ndtr(f)
        pdf = np.exp(-0.5 * f * f) / np.sqrt(2 * np.pi)
        a, b = p + 1e-5, 1 - p + 1e-5
        l = ys * np.log(a) + (1 - ys) * np.log(b)
        d1 = ys * pdf / a - (1 - ys) * pdf / b
        dpdf = -f * pdf
        d2 = ys * (dpdf / a - pdf * pdf / (a * a)) - (1 - ys) * (dpdf / b + pdf * pdf / (b * b))
    else:
        raise ValueError(kind)
    z = lambda q: np.where(obs, q @ w, 0.0)   # noqa: E731
    return z(l), z(d1), 0.5 * z(d2)


def surrogate_posterior(prior_args, t, Ytil, Vtil, jitter, nthreads=0):
    """q(u) marginals + lml of the surrogate SDE_GP whose data / noise are the sites.
    prior_args = (block_size, lam [B, nblk], Pinf [B, d, d], H [1, d]); Ytil, Vtil [B, T]."""
    bs, lam, Pinf, H = prior_args
    out = c_oracle.filter_smooth(bs, lam, Pinf, H, t, Ytil[..., None], Vtil[..., None, None], jitter=jitter,
                                 full_state=False, keep_filtered=False, nthreads=nthreads)
    return out["lml"], out["ms"][..., 0], out["Ps"][..., 0, 0], out["threads"]


def natural_gradient_update(prior_args, t, Y, Ytil, Vtil, kind, beta, K=20, binsize=1.0, jitter=la.JITTER,
                            ng_jitter=la.NG_JITTER, nthreads=0):
    """cvi_nat_grad.py:346-410 + cvi_parameterisations.py:63-93 for D = 1.  Returns new (Ytil, Vtil)."""
    _, q_mu, q_var, _ = surrogate_posterior(prior_args, t, Ytil, Vtil, jitter, nthreads)
    _, dm, dS = gh_ell_and_grads(Y, q_mu, q_var, kind, K, binsize)
    vj = Vtil + ng_jitter                                  # theta_to_lambda
    l1, l2 = Ytil / vj, -0.5 / vj
    g1 = dm - 2.0 * dS * q_mu                              # cvi_block_update
    l1n, l2n = (1 - beta) * l1 + beta * g1, (1 - beta) * l2 + beta * dS
    th2 = 1.0 / (-2.0 * l2n + ng_jitter)                   # lambda_to_theta
    return th2 * l1n, th2


def elbo(prior_args, t, Y, Ytil, Vtil, kind, K=20, binsize=1.0, jitter=la.JITTER, nthreads=0):
    """elbos.py:163-194 per block: ELL(data) - ELL(surrogate) + lml(surrogate)."""
    lml, q_mu, q_var, _ = surrogate_posterior(prior_args, t, Ytil, Vtil, jitter, nthreads)
    ell, _, _ = gh_ell_and_grads(Y, q_mu, q_var, kind, K, binsize)
    # full_gaussian_expected_log_likelihood (expected_log_likelihoods.py:90-117), scalar, nothing missing
    ell_s = -0.5 * (LOG2PI + np.log(Vtil) + (Ytil - q_mu) ** 2 / Vtil) - 0.5 * q_var / Vtil
    return ell.sum(-1) - ell_s.sum(-1) + lml


def cvi_iteration(prior_args, t, Y, Ytil, Vtil, kind, beta, **kw):
    """One bench 'step': natural_gradient_update then the ELBO (vgp.py:274-282, 148-157)."""
    Yn, Vn = natural_gradient_update(prior_args, t, Y, Ytil, Vtil, kind, beta, **kw)
    kw.pop("ng_jitter", None)
    return Yn, Vn, elbo(prior_args, t, Y, Yn, Vn, kind, **kw)
