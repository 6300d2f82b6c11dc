"""Oracle: reverse-mode derivative (vector-Jacobian product) of the sequential Kalman filter's log marginal
likelihood -- TEST INFRASTRUCTURE, never imported by the product.

What the reference computes with `jax.jacrev` / `jax.grad` THROUGH `filter('sequential')` for hyper-parameter
steps (stgp/trainers/trainer.py:43,128-136; stgp/trainers/standard.py:58-91).  The
reference holds no hand-written adjoint: this file restates the chain rule of the forward recursion in
oracle/filters.py (kalman_filter.py:144-241,439-485), step by step in reverse, and is pinned by
  (a) torch autograd through a torch transcription of the same forward recursion (tests/test_oracle_adjoint.py),
  (b) central finite differences of oracle/filters.py.

Forward step k (see oracle/filters.py):
    m_ = A m ;  P_ = A P A^T + Q ;  Hm = M H ;  v = y0 - Hm m_ ;  S = Hm P_ Hm^T + R ;  Sj = S + jitter I
    X = Sj^-1 (Hm P_) ;  K = X^T ;  m+ = m_ + K v ;  P+ = P_ - K S K^T
    l_k = log N(y0 | Hm m_, mask_to_identity(S)) restricted to the observed entries
"""
import numpy as np

from . import linalg as la


def _update_vjp(m_, P_, H, R, y, jitter, mbar, Pbar, gbar):
    """VJP of kf_update_step.  Returns (m_bar_, P_bar_, Hbar, Rbar)."""
    mdim = y.shape[0]
    mask = (~np.isnan(y[:, 0]))
    y0 = np.nan_to_num(y)
    M = np.diag(mask.astype(float))
    Hm = M @ H
    v = y0 - Hm @ m_
    W = Hm @ P_
    S = W @ Hm.T + R
    Sj = S + jitter * np.eye(mdim)
    Sj_inv = np.linalg.inv(Sj)
    X = Sj_inv @ W
    K = X.T
    # lml term: un-jittered S, missing rows / cols -> identity
    Sm = la.mask_to_identity(S, mask.astype(int))
    Sm_inv = np.linalg.inv(Sm)
    a = Sm_inv @ v
    mm = np.outer(mask, mask).astype(float)

    Kbar = mbar @ v.T - (Pbar @ K @ S.T + Pbar.T @ K @ S)
    Sbar = -K.T @ Pbar @ K
    vbar = K.T @ mbar - gbar * a
    Sbar = Sbar + gbar * (-0.5) * (Sm_inv - a @ a.T) * mm
    Xbar = Kbar.T
    Wbar = Sj_inv.T @ Xbar
    Sbar = Sbar - Sj_inv.T @ Xbar @ X.T
    Rbar = Sbar.copy()
    Hmbar = (Sbar + Sbar.T) @ Hm @ P_ + Wbar @ P_.T - vbar @ m_.T
    P_bar = Pbar + Hm.T @ Sbar @ Hm + Hm.T @ Wbar
    m_bar = mbar - Hm.T @ vbar
    Hbar = M @ Hmbar
    return m_bar, P_bar, Hbar, Rbar


def filter_lml_vjp(A, Q, H, R, Y, m0, P0, jitter=la.JITTER, gbar=1.0):
    """Gradient of lml = sum_k l_k of the sequential filter driven by given transitions.

    A, Q [T, d, d]; H [m, d]; R [T, m, m]; Y [T, m] (NaN = missing); m0 [d, 1]; P0 [d, d].
    Returns dict(lml, gA [T,d,d], gQ [T,d,d], gH [m,d], gR [T,m,m], gm0 [d,1], gP0 [d,d]).
    Every matrix entry is treated as an independent variable (no symmetry assumed)."""
    T = A.shape[0]
    Ycol = np.reshape(Y, [T, -1])[..., None]
    # forward, keeping the state BEFORE each step
    from .filters import kf_update_step
    ms, Ps = [m0], [P0]
    lml = 0.0
    m, P = m0, P0
    for k in range(T):
        m_ = A[k] @ m
        P_ = A[k] @ P @ A[k].T + Q[k]
        m, P, l = kf_update_step(m_, P_, H, R[k], Ycol[k], jitter)
        lml += l
        ms.append(m), Ps.append(P)
    d = m0.shape[0]
    mbar, Pbar = np.zeros((d, 1)), np.zeros((d, d))
    gA, gQ = np.zeros_like(A), np.zeros_like(Q)
    gR = np.zeros_like(R)
    gH = np.zeros_like(H)
    for k in range(T - 1, -1, -1):
        mp, Pp = ms[k], Ps[k]
        m_ = A[k] @ mp
        P_ = A[k] @ Pp @ A[k].T + Q[k]
        m_bar, P_bar, Hb, Rb = _update_vjp(m_, P_, H, R[k], Ycol[k], jitter, mbar, Pbar, gbar)
        gH += Hb
        gR[k] = Rb
        gQ[k] = P_bar
        gA[k] = m_bar @ mp.T + P_bar @ A[k] @ Pp.T + P_bar.T @ A[k] @ Pp
        mbar = A[k].T @ m_bar
        Pbar = A[k].T @ P_bar @ A[k]
    return dict(lml=float(lml), gA=gA, gQ=gQ, gH=gH, gR=gR, gm0=mbar, gP0=Pbar)


def matern_chain(prior_blocks, lam, dt, Pinf, gA, gQ, dA_dlam):
    """Chain (gA, gQ) of a stationary block-diagonal Matern prior to (glam [nblk], gPinf [d, d]):
       Q_k = Pinf - A_k Pinf A_k^T  (kernels/kernel.py:207-209),  A_k = blockdiag(expm(F(lam_b) dt_k)).
    dA_dlam(b, dt) -> [s, s] derivative of block b's transition with respect to lam_b."""
    T, d, _ = gA.shape
    nblk = len(lam)
    s = d // nblk
    glam = np.zeros(nblk)
    gPinf = np.zeros((d, d))
    for k in range(T):
        Ak = prior_blocks(dt[k])
        Qb = gQ[k]
        gPinf += Qb - Ak.T @ Qb @ Ak
        Abar = gA[k] - (Qb @ Ak @ Pinf.T + Qb.T @ Ak @ Pinf)
        for b in range(nblk):
            sl = slice(b * s, (b + 1) * s)
            glam[b] += np.sum(Abar[sl, sl] * dA_dlam(b, dt[k]))
    return glam, gPinf
