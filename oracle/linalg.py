"""Oracle: small dense helpers with the reference's jitter placement (test infrastructure).

Follows (paths relative to /root/reference/src/lib/stgp/):
  computation/linalg.py:12-33            solve(): ALWAYS adds settings.jitter*I, then Cholesky solve
  computation/matrix_ops.py:107-110      add_jitter
  computation/matrix_ops.py:220-236      cholesky (lower), cholesky_solve
  computation/matrix_ops.py:382-385      mat_inv (jittered)
  computation/matrix_ops.py:405-407      force_symmetric
  computation/gaussian.py:42-108         log_gaussian (no jitter), log_gaussian_with_mask
  utils/nan_utils.py:13-69               get_same_shape_mask, mask_vector, mask_to_identity
  settings.py:63-64                      jitter = 1e-5, ng_jitter = 1e-7
"""
import numpy as np
import scipy.linalg as sla

JITTER = 1e-5      # settings.py:63
NG_JITTER = 1e-7   # settings.py:64
LOG2PI = np.log(2 * np.pi)


def cholesky(A):
    """Lower Cholesky factor.  A non-PD input gives NaNs (JAX semantics), never an exception."""
    try:
        return np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        return np.full_like(A, np.nan)


def cholesky_solve(L, X):
    if np.any(np.isnan(L)):
        return np.full((L.shape[0],) + X.shape[1:], np.nan)
    return sla.cho_solve((L, True), X)


def add_jitter(A, jit):
    return A + jit * np.eye(A.shape[0])


def solve(A, B, jitter=JITTER):
    """linalg.py:12-33 with linear_solver == CHOLESKY (the default, settings.py:46)."""
    return cholesky_solve(cholesky(add_jitter(A, jitter)), B)


def mat_inv(A, jitter=JITTER):
    return cholesky_solve(cholesky(add_jitter(A, jitter)), np.eye(A.shape[0]))


def force_symmetric(A):
    return 0.5 * (A + A.T)


def mask_to_identity(K, mask):
    """nan_utils.py:49-69 -- rows/cols of missing entries replaced by identity rows/cols."""
    N = K.shape[0]
    mm = np.tile(mask, [N, 1])
    K = K - np.eye(N)
    K = K * mm
    K = K * mm.T
    return K + np.eye(N)


def log_gaussian(Y, mu, sigma):
    """gaussian.py:42-69 (Cholesky of the UN-jittered sigma)."""
    L = cholesky(sigma)
    N = Y.shape[0]
    c = -0.5 * N * LOG2PI - 0.5 * np.sum(np.log(np.square(np.diag(L))))
    err = Y - mu
    mahal = err.T @ cholesky_solve(L, err)
    return float(np.squeeze(c - 0.5 * mahal))


def log_gaussian_with_mask(Y, mu, sigma, mask):
    """gaussian.py:72-108."""
    Y = np.nan_to_num(Y, nan=0.0)
    sigma = mask_to_identity(sigma, mask)
    mu = np.where(mask[:, None].astype(bool), mu, 0.0)
    N_mask = np.sum(1 - mask)
    return log_gaussian(Y, mu, sigma) + 0.5 * N_mask * LOG2PI
