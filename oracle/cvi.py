"""Oracle: CVI natural-gradient site update, expected log-likelihoods and ELBO (test infrastructure).

Follows (paths relative to /root/reference/src/lib/stgp/):
  computation/natural_gradients/exponential_family_transforms.py:25-42   theta_to_lambda   (ng_jitter)
  computation/natural_gradients/exponential_family_transforms.py:70-83   lambda_to_theta   (ng_jitter)
  computation/natural_gradients/cvi_nat_grad.py:47-87                    cvi_block_update
  computation/natural_gradients/cvi_nat_grad.py:346-410                  natural_gradients(FullConjugateGaussian)
  computation/natural_gradients/cvi_parameterisations.py:63-93           lambda -> theta re-entry ('NG_Moment')
  computation/elbos/expected_log_likelihoods.py:90-117                   full_gaussian_expected_log_likelihood
  computation/elbos/elbos.py:163-194                                     elbo(FullConjugateGaussian)
  computation/general.py:9-26, likelihood/poisson.py:20-22, likelihood/bernoulli.py:17-19   scalar log-liks
  computation/integrals/samples.py:67-90                                 Gauss-Hermite branch (DEAD code in
       the reference: undefined names, use_quadrature=False) -- restated here as the formula it intends.

PARITY STATUS.  theta<->lambda, cvi_block_update, the Gaussian/surrogate ELL and the ELBO assembly are
deterministic in the reference and restated verbatim.  The reference obtains dELL/dm, dELL/dS by
`jax.grad` (cvi_nat_grad.py:381-383); the closed forms used here (Gaussian: R^-1(y-m), -1/2 R^-1) are
checked against finite differences of the restated ELL in tests/test_oracle_cvi.py.  The non-Gaussian
ELL in the reference is Monte-Carlo on an objax PRNG stream (not reproducible without JAX); its
Gauss-Hermite replacement below is **parity unpinned** against the reference and is validated against
the Poisson closed form (expected_log_likelihoods.py:149-174) and high-order quadrature instead.
"""
import numpy as np
from scipy.special import gammaln, ndtr

from . import linalg as la


# ------------------------------------------------------------------ exponential-family transforms

def theta_to_lambda(theta_1, theta_2, ng_jitter=la.NG_JITTER):
    """theta = (Y~ [D,1], V~ [D,D]) -> lambda = (V~^-1 Y~, -1/2 V~^-1)."""
    D = theta_1.shape[0]
    L = la.cholesky(theta_2 + ng_jitter * np.eye(D))
    lambda_1 = la.cholesky_solve(L, theta_1)
    lambda_2 = -0.5 * la.cholesky_solve(L, np.eye(D))
    return lambda_1, lambda_2


def lambda_to_theta(lambda_1, lambda_2, ng_jitter=la.NG_JITTER):
    D = lambda_1.shape[0]
    L = la.cholesky(la.add_jitter(-2 * lambda_2, ng_jitter))
    theta_2 = la.cholesky_solve(L, np.eye(D))
    theta_1 = theta_2 @ lambda_1
    return theta_1, theta_2


def theta_precision_to_lambda(theta_1, theta_2, ng_jitter=la.NG_JITTER):
    """exponential_family_transforms.py:44-53 ('NG_Precision': theta_2 is the site PRECISION): lambda_2 = -1/2 theta_2
    and -- exactly as the reference writes it -- lambda_1 = (theta_2 + ng_jitter I)^-1 theta_1, a cholesky_solve with
    the precision's factor (not the product precision @ theta_1)."""
    D = theta_1.shape[0]
    L = la.cholesky(theta_2 + ng_jitter * np.eye(D))
    return la.cholesky_solve(L, theta_1), -0.5 * theta_2


def lambda_to_theta_precision(lambda_1, lambda_2, ng_jitter=la.NG_JITTER):
    """exponential_family_transforms.py:85-95: theta_2 = -2 lambda_2 (precision), theta_1 = (-2 lambda_2 + jit I)^-1 lambda_1."""
    D = lambda_1.shape[0]
    L = la.cholesky(-2 * lambda_2 + ng_jitter * np.eye(D))
    return la.cholesky_solve(L, lambda_1), -2 * lambda_2


def cvi_block_update(lambda_1, lambda_2, m, s, m_grad, s_grad, beta):
    """cvi_nat_grad.py:47-87 with enforce_psd_type=None."""
    grad_1 = m_grad - 2 * s_grad @ m
    grad_2 = s_grad
    return (1 - beta) * lambda_1 + beta * grad_1, (1 - beta) * lambda_2 + beta * grad_2


# ------------------------------------------------------------------ expected log-likelihoods

def full_gaussian_ell(Y, noise, q_mu, q_covar):
    """expected_log_likelihoods.py:90-117: log N(Y | q_mu, noise) - 1/2 tr(noise^-1 q_covar), NaN-masked.
    Y, q_mu [P,1]; noise, q_covar [P,P]."""
    mask = (~np.isnan(Y[:, 0])).astype(int)
    ml = la.log_gaussian_with_mask(Y, q_mu, noise, mask)
    N_mask = np.sum(1 - mask)
    masked_noise = la.mask_to_identity(noise, mask)
    L = la.cholesky(masked_noise)
    masked_q = la.mask_to_identity(q_covar, mask)
    traced = la.cholesky_solve(L, masked_q)
    return ml - 0.5 * np.trace(traced) + 0.5 * N_mask


def gaussian_ell_and_grads(Y, noise, W, q_mu, q_var):
    """ELL of a Gaussian likelihood y = W u + e, e ~ N(0, noise), under q(u) = N(q_mu, q_var), with the
    gradients w.r.t. (q_mu, q_var) that jax.grad produces in cvi_nat_grad.py:381-383.
    W=None means W = I.  Y [P]; noise [P,P]; q_mu [D]; q_var [D,D]."""
    D = q_mu.shape[0]
    Wm = np.eye(D) if W is None else W
    f_mu = Wm @ q_mu
    f_var = Wm @ q_var @ Wm.T
    ell = full_gaussian_ell(Y[:, None], noise, f_mu[:, None], f_var)
    mask = ~np.isnan(Y)
    Rinv = np.zeros_like(noise)
    if mask.any():
        idx = np.where(mask)[0]
        Rinv[np.ix_(idx, idx)] = np.linalg.inv(noise[np.ix_(idx, idx)])
    err = np.where(mask, np.nan_to_num(Y) - f_mu, 0.0)
    dm = Wm.T @ (Rinv @ err)
    dS = -0.5 * Wm.T @ Rinv @ Wm
    return ell, dm, dS


def spatial_sparsity_gaussian_ell_and_grads(Y, s2, W, c0_diag, jitter, q_mu, q_var):
    """ELL of a Gaussian likelihood with noise variance s2 at N data locations whose marginals come from the block
    posterior q(u) = N(q_mu, q_var) at M inducing locations through the spatial conditional
    (computation/spatial_conditionals.py:30-207 -> marginals.py:82-113):
        m_x = W q_mu,   v_x = c0_diag + diag(W (q_var + jitter I) W^T),   ELL = sum_obs log N(y | m_x, s2) - v_x / (2 s2)
    with the gradients w.r.t. (q_mu, q_var) that `jax.grad(partial_ell, (1, 2))` (cvi_nat_grad.py:381-383) produces for
    the SpatialSparsity posterior.  Y [N] (NaN = missing), W [N, M], c0_diag [N]."""
    M = q_mu.shape[0]
    m_x = W @ q_mu
    v_x = c0_diag + np.einsum("ij,jk,ik->i", W, q_var + jitter * np.eye(M), W)
    obs = ~np.isnan(Y)
    r = np.where(obs, np.nan_to_num(Y) - m_x, 0.0)
    ell = float(np.sum(np.where(obs, -0.5 * (np.log(2 * np.pi * s2) + (r * r + v_x) / s2), 0.0)))
    dm = W.T @ (r / s2)
    dS = -0.5 * (W.T * np.where(obs, 1.0 / s2, 0.0)) @ W
    return ell, dm, dS


def spatial_sparsity_gh_ell_and_grads(Y, kind, W, c0_diag, jitter, q_mu, q_var, K=20, binsize=1.0):
    """The same pull-back for independent non-Gaussian observations at the N data locations (Poisson exp-link /
    Bernoulli probit): per point the K-point Gauss-Hermite E[l], E[l'], 1/2 E[l''] under N(m_x,i, v_x,i), then
    dELL/dq_mu = W^T E[l'],  dELL/dq_var = W^T diag(1/2 E[l'']) W   (chain rule through m_x = W q_mu and
    v_x = c0 + diag(W (q_var + jitter I) W^T); what jax.grad of the reference's ELL gives, cvi_nat_grad.py:381-383)."""
    M = q_mu.shape[0]
    m_x = W @ q_mu
    v_x = c0_diag + np.einsum("ij,jk,ik->i", W, q_var + jitter * np.eye(M), W)
    g = [gh_ell_and_grads(Y[i], m_x[i], v_x[i], kind, K, binsize) for i in range(Y.shape[0])]
    e0, e1, e2 = (np.array([x[j] for x in g]) for j in range(3))
    return float(e0.sum()), W.T @ e1, (W.T * e2) @ W


def log_poisson(y, f, binsize=1.0):
    """general.py:9-11 with likelihood/poisson.py:20-22 (exp link)."""
    lam = np.exp(f) * binsize
    return y * np.log(lam) - lam - gammaln(y + 1.0)


def log_bernoulli_probit(y, f):
    """general.py:13-26 with likelihood/bernoulli.py:17-19: probit link, +1e-5 inside both logs."""
    p = ndtr(f)
    return y * np.log(p + 1e-5) + (1 - y) * np.log(1 - p + 1e-5)


def _dlog_poisson(y, f, binsize):
    lam = np.exp(f) * binsize
    return y - lam, -lam


def _dlog_bernoulli(y, f):
    p = ndtr(f)
    pdf = np.exp(-0.5 * f * f) / np.sqrt(2 * np.pi)
    a, b = p + 1e-5, 1 - p + 1e-5
    d1 = y * pdf / a - (1 - y) * pdf / b
    dpdf = -f * pdf
    d2 = y * (dpdf / a - pdf * pdf / (a * a)) - (1 - y) * (dpdf / b + pdf * pdf / (b * b))
    return d1, d2


def gh_ell_and_grads(y, m, v, kind, K=20, binsize=1.0):
    """K-point Gauss-Hermite expected log-likelihood of a scalar site and its derivatives w.r.t. the
    marginal mean m and variance v:  E[l(f)], E[l'(f)], 1/2 E[l''(f)]   (f ~ N(m, v)).
    NaN y -> (0, 0, 0).  PARITY UNPINNED (see module docstring)."""
    if np.isnan(y):
        return 0.0, 0.0, 0.0
    x, w = np.polynomial.hermite.hermgauss(K)
    w = w / np.sqrt(np.pi)
    f = m + np.sqrt(2.0 * v) * x
    if kind == "poisson":
        l = log_poisson(y, f, binsize)
        d1, d2 = _dlog_poisson(y, f, binsize)
    elif kind == "bernoulli":
        l = log_bernoulli_probit(y, f)
        d1, d2 = _dlog_bernoulli(y, f)
    else:
        raise ValueError(kind)
    return float(np.sum(w * l)), float(np.sum(w * d1)), float(0.5 * np.sum(w * d2))


def poisson_ell_closed_form(y, m, v, binsize=1.0):
    """expected_log_likelihoods.py:149-174."""
    return y * np.log(binsize) + y * m - binsize * np.exp(m + v / 2) - gammaln(y + 1.0)


# ------------------------------------------------------------------ one CVI iteration / ELBO

def cvi_step(Ytil, Vtil, q_mu, q_var, dm, dS, beta, ng_jitter=la.NG_JITTER):
    """natural_gradients(FullConjugateGaussian) + 'NG_Moment' re-entry for all T blocks.
    Ytil [T,D]; Vtil, q_var, dS [T,D,D]; q_mu, dm [T,D].  Returns new (Ytil, Vtil)."""
    T, D = Ytil.shape
    Yn, Vn = np.empty_like(Ytil), np.empty_like(Vtil)
    for t in range(T):
        l1, l2 = theta_to_lambda(Ytil[t][:, None], Vtil[t], ng_jitter)
        l1n, l2n = cvi_block_update(l1, l2, q_mu[t][:, None], q_var[t], dm[t][:, None], dS[t], beta)
        th1, th2 = lambda_to_theta(l1n, l2n, ng_jitter)
        Yn[t], Vn[t] = th1[:, 0], th2
    return Yn, Vn


def cvi_step_precision(Ytil, Ptil, q_mu, q_var, dm, dS, beta, ng_jitter=la.NG_JITTER):
    """natural_gradients(FullConjugateGaussian) + the 'NG_Precision' re-entry (cvi_parameterisations.py:95-113,
    cvi_nat_grad_utils.py:62-63) for all T blocks: the sites carry (Y~, precision).  Returns new (Ytil, Ptil)."""
    T, D = Ytil.shape
    Yn, Pn = np.empty_like(Ytil), np.empty_like(Ptil)
    for t in range(T):
        l1, l2 = theta_precision_to_lambda(Ytil[t][:, None], Ptil[t], ng_jitter)
        l1n, l2n = cvi_block_update(l1, l2, q_mu[t][:, None], q_var[t], dm[t][:, None], dS[t], beta)
        th1, th2 = lambda_to_theta_precision(l1n, l2n, ng_jitter)
        Yn[t], Pn[t] = th1[:, 0], th2
    return Yn, Pn


def surrogate_ell(Ytil, Vtil, q_mu, q_var):
    """ELL of the surrogate likelihood N(Y~ | u, V~) under q (elbos.py:181-188)."""
    return float(sum(full_gaussian_ell(Ytil[t][:, None], Vtil[t], q_mu[t][:, None], q_var[t])
                     for t in range(Ytil.shape[0])))


def elbo(ell_data, ell_surrogate, lml_surrogate):
    """elbos.py:189: ELBO = ELL - ELL_surrogate + ML_surrogate."""
    return ell_data - ell_surrogate + lml_surrogate


# ------------------------------------------------------------------ damped-oscillator collocation (config 3)

def pendulum_forward(u, g_over_l, damping):
    """transforms/pdes.py:584-597 (DampedPendulum1D.forward) stacked under the observation map of
    zoo/sde_diff.py:757-763: T(u) = [x, x_tt + (g/l) sin x + b x_t] for u = (x, x_t, x_tt)."""
    return np.array([u[0], u[2] + g_over_l * np.sin(u[0]) + damping * u[1]])


def pendulum_ell_and_grads(y, q_mu, q_var, g_over_l, damping, var_obs, var_col, gauss_newton=False, idx=(0, 1, 2)):
    """Expected log-likelihood of y = (observation of x, collocation target) under independent Gaussian
    noise (var_obs, var_col) on T(u), u ~ N(q_mu, q_var), with its gradients w.r.t. (q_mu, q_var).

    The reference evaluates this expectation by Monte-Carlo (approximators.py:16-58) and differentiates it
    with jax.grad (cvi_nat_grad.py:381-383); here it is closed-form (Gaussian moments of sin / cos + Stein's
    lemma), checked against dense Gauss-Hermite quadrature and finite differences in tests/test_oracle_cvi.py.
    gauss_newton=True replaces dELL/dS by 1/2 sum_p mask_p J_p^T (-1/var_p) J_p, J_p = dT_p/du at u = q_mu
    (cvi_hessian_approximations.py:333-431,483-486: Laplace Gauss-Newton, delta-u, delta-f)."""
    i0, i1, i2 = idx
    D = q_mu.shape[0]
    a, b = g_over_l, damping
    m0, m1, m2 = q_mu[i0], q_mu[i1], q_mu[i2]
    S = 0.5 * (q_var + q_var.T)
    ell, dm, dS = 0.0, np.zeros(D), np.zeros((D, D))
    if not np.isnan(y[0]):
        e = y[0] - m0
        ell += -0.5 * (np.log(2 * np.pi) + np.log(var_obs) + (e * e + S[i0, i0]) / var_obs)
        dm[i0] += e / var_obs
        dS[i0, i0] += -0.5 / var_obs
    if not np.isnan(y[1]):
        E, E4 = np.exp(-0.5 * S[i0, i0]), np.exp(-2.0 * S[i0, i0])
        sm, cm = np.sin(m0), np.cos(m0)
        mul = m2 + b * m1 - y[1]
        vl = S[i2, i2] + 2 * b * S[i1, i2] + b * b * S[i1, i1]
        c = S[i0, i2] + b * S[i0, i1]
        Rr = mul * mul + vl + 2 * a * (mul * sm + c * cm) * E + a * a * 0.5 * (1 - np.cos(2 * m0) * E4)
        ell += -0.5 * (np.log(2 * np.pi) + np.log(var_col) + Rr / var_col)
        k = -0.5 / var_col
        t = 2 * mul + 2 * a * sm * E
        dm[i0] += k * (2 * a * (mul * cm - c * sm) * E + a * a * np.sin(2 * m0) * E4)
        dm[i1] += k * b * t
        dm[i2] += k * t
        if gauss_newton:
            J = np.zeros(D)
            J[i0], J[i1], J[i2] = a * cm, b, 1.0
            dS += k * np.outer(J, J)
        else:
            G = np.zeros((D, D))
            G[i0, i0] = -a * (mul * sm + c * cm) * E + a * a * np.cos(2 * m0) * E4
            G[i1, i1], G[i2, i2] = b * b, 1.0
            G[i0, i1] = G[i1, i0] = a * b * cm * E
            G[i0, i2] = G[i2, i0] = a * cm * E
            G[i1, i2] = G[i2, i1] = b
            dS += k * G
    return float(ell), dm, dS


def pendulum_ell_quadrature(y, q_mu, q_var, g_over_l, damping, var_obs, var_col, K=40):
    """The same expectation by K^3-point tensor Gauss-Hermite over (x, x_t, x_tt): independent check."""
    x, w = np.polynomial.hermite.hermgauss(K)
    w = w / np.sqrt(np.pi)
    L = np.linalg.cholesky(q_var[:3, :3] + 1e-300 * np.eye(3))
    Z = np.stack(np.meshgrid(x, x, x, indexing="ij"), -1).reshape(-1, 3) * np.sqrt(2.0)
    Wt = np.einsum("i,j,k->ijk", w, w, w).reshape(-1)
    U = q_mu[:3] + Z @ L.T
    ell = 0.0
    if not np.isnan(y[0]):
        ell += np.sum(Wt * (-0.5 * (np.log(2 * np.pi * var_obs) + (y[0] - U[:, 0]) ** 2 / var_obs)))
    if not np.isnan(y[1]):
        r = U[:, 2] + g_over_l * np.sin(U[:, 0]) + damping * U[:, 1]
        ell += np.sum(Wt * (-0.5 * (np.log(2 * np.pi * var_col) + (y[1] - r) ** 2 / var_col)))
    return float(ell)
