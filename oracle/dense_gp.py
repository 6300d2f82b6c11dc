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
