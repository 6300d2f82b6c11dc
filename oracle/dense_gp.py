"""Oracle cross-check: dense O(N^3) GP regression (test infrastructure).

The independent known-answer the reference supports by construction:
  computation/log_marginal_likelihoods.py:36-58  log N(y | 0, K + sigma^2 I)
  computation/gaussian.py:42-69                  log_gaussian
with K from the Matern kernels (kernels/matern.py:82-90,179-188,331-341).
"""
import numpy as np
import scipy.linalg as sla


def log_marginal_likelihood(K, y, noise_var):
    N = y.shape[0]
    Ky = K + noise_var * np.eye(N)
    L = np.linalg.cholesky(Ky)
    alpha = sla.cho_solve((L, True), y)
    return float(-0.5 * N * np.log(2 * np.pi) - np.sum(np.log(np.diag(L))) - 0.5 * y @ alpha)


def posterior(K, y, noise_var):
    """Posterior mean and marginal variance of f at the training inputs."""
    N = y.shape[0]
    Ky = K + noise_var * np.eye(N)
    L = np.linalg.cholesky(Ky)
    mean = K @ sla.cho_solve((L, True), y)
    V = sla.solve_triangular(L, K, lower=True)
    var = np.diag(K) - np.sum(V * V, axis=0)
    return mean, var


def spatial_conditional(Kzz, Ksz, Kss, Ktt, pred_mean, pred_var, jitter):
    """Posterior at new spatial points from the per-step posterior at the inducing points -- the `f_only` branch of
    `spatial_conditional_block` (computation/spatial_conditionals.py:137-207: per-step Cholesky of
    pred_var + jitter I, Cholesky of Kzz + jitter I, then the vmapped `gaussian_spatial_conditional_cholesky`,
    computation/marginals.py:82-113), restated step by step:
        A = L_zz^-1 Kzs, A1 = L_zz^-T A, A2 = (S_chol^T A1)^T,  mu = A1^T m,  sig = Ktt (Kss - A^T A) + A2 A2^T.
    Kzz [M, M], Ksz [N, M], Kss [N, N], Ktt [T] (variance of the temporal kernel per step), pred_mean [T, M, 1],
    pred_var [T, M, M].  Returns mu [T, N, 1], var [T, 1, N, N] (:198-201)."""
    Kzz, Ksz, Kss = (np.asarray(x, float) for x in (Kzz, Ksz, Kss))
    M = Kzz.shape[0]
    Lzz = np.linalg.cholesky(Kzz + jitter * np.eye(M))
    A = sla.solve_triangular(Lzz, Ksz.T, lower=True)
    A1 = sla.solve_triangular(Lzz.T, A, lower=False)
    C0 = Kss - A.T @ A
    T = pred_mean.shape[0]
    N = Kss.shape[0]
    mu = np.zeros([T, N, 1])
    var = np.zeros([T, 1, N, N])
    for t in range(T):
        S_chol = np.linalg.cholesky(pred_var[t] + jitter * np.eye(M))
        A2 = (S_chol.T @ A1).T
        mu[t] = A1.T @ pred_mean[t].reshape(M, 1)
        var[t, 0] = Ktt[t] * C0 + A2 @ A2.T
    return mu, var
