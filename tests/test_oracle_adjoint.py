"""Pins for oracle/adjoint.py: the hand-derived reverse pass against torch
autograd through a torch transcription of the reference's forward recursion -- what `jax.jacrev` through
`filter('sequential')` computes (stgp/trainers/trainer.py:128-136) -- and against central differences of
oracle/filters.py."""
import math

import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import adjoint, sde
from oracle import filters as ofilters


def _torch_filter_lml(A, Q, H, R, Y, m0, P0, jitter):
    """kalman_filter.py:144-241,439-485 in torch (fp64), differentiable.  Y numpy [T, m] with NaN."""
    T, mdim = Y.shape
    m, P = m0, P0
    lml = torch.zeros((), dtype=torch.float64)
    for k in range(T):
        mask = torch.tensor(~np.isnan(Y[k]), dtype=torch.float64)
        y0 = torch.tensor(np.nan_to_num(Y[k]), dtype=torch.float64)[:, None]
        M = torch.diag(mask)
        m_ = A[k] @ m
        P_ = A[k] @ P @ A[k].T + Q[k]
        Hm = M @ H
        v = y0 - Hm @ m_
        S = Hm @ P_ @ Hm.T + R[k]
        Sj = S + jitter * torch.eye(mdim, dtype=torch.float64)
        K = torch.linalg.solve(Sj, Hm @ P_).T
        m = m_ + K @ v
        P = P_ - K @ S @ K.T
        Sm = S * torch.outer(mask, mask) + torch.diag(1.0 - mask)
        nobs = mask.sum()
        lml = lml - 0.5 * (nobs * math.log(2 * math.pi) + torch.logdet(Sm) + (v.T @ torch.linalg.solve(Sm, v))[0, 0])
    return lml


def _problem(seed, T=14, d=3, m=2):
    rng = np.random.default_rng(seed)
    A = 0.8 * np.eye(d)[None] + 0.1 * rng.normal(size=(T, d, d))
    L = rng.normal(size=(T, d, d)) * 0.3
    Q = L @ np.transpose(L, (0, 2, 1)) + 0.05 * np.eye(d)
    H = rng.normal(size=(m, d))
    Lr = rng.normal(size=(T, m, m)) * 0.3
    R = Lr @ np.transpose(Lr, (0, 2, 1)) + 0.2 * np.eye(m)
    Y = rng.normal(size=(T, m))
    Y[rng.uniform(size=(T, m)) < 0.25] = np.nan
    Y[3] = np.nan                                             # a fully missing step
    m0 = rng.normal(size=(d, 1))
    L0 = rng.normal(size=(d, d))
    P0 = L0 @ L0.T + 0.5 * np.eye(d)
    return A, Q, H, R, Y, m0, P0


@pytest.mark.parametrize("jitter", [1e-5, 0.0, 1e-2])
@pytest.mark.parametrize("seed", [0, 1])
def test_adjoint_matches_autograd(seed, jitter):
    A, Q, H, R, Y, m0, P0 = _problem(seed)
    tv = [torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (A, Q, H, R, m0, P0)]
    lml = _torch_filter_lml(tv[0], tv[1], tv[2], tv[3], Y, tv[4], tv[5], jitter)
    lml.backward()
    g = adjoint.filter_lml_vjp(A, Q, H, R, Y, m0, P0, jitter=jitter)
    assert abs(g["lml"] - float(lml.detach())) <= 1e-12 * abs(float(lml.detach()))
    for name, t in zip(("gA", "gQ", "gH", "gR", "gm0", "gP0"), tv):
        ref = t.grad.numpy()
        assert np.abs(g[name] - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-300), name


def test_adjoint_matches_finite_differences_of_the_oracle_filter():
    """Directional derivative of oracle/filters.py's own lml along a random direction of the hyper-parameters
    (lengthscale via lam, variance via Pinf, noise) of a Matern-5/2 + Matern-3/2 sum."""
    rng = np.random.default_rng(5)
    T = 40
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    y = rng.normal(size=(T, 1))
    y[rng.uniform(size=T) < 0.1] = np.nan

    def lml_of(ls1, ls2, noise):
        prior = sde.LTI_SDE([sde.SumKernel([sde.Matern52(ls1, 0.7), sde.Matern32(ls2, 0.4)])])
        R = np.tile(np.array([[noise]]), [T, 1, 1])
        return ofilters.filter_sequential(prior, t, y, R, jitter=1e-5)[0], prior

    th = np.array([0.8, 0.5, 0.2])
    _, prior = lml_of(*th)
    dt = np.hstack([0.0, np.diff(t)])
    Pinf = prior.P_inf()
    A = np.array([prior.expm(x) for x in dt])
    Q = np.array([prior.Q(x, a, Pinf) for x, a in zip(dt, A)])
    R = np.tile(np.array([[th[2]]]), [T, 1, 1])
    g = adjoint.filter_lml_vjp(A, Q, prior.H(), R, y, prior.m_inf(), Pinf, jitter=1e-5)
    # chain to (ls1, ls2, noise) by finite differences of the (cheap, T-independent) prior quantities
    eps = 1e-6
    grad = np.zeros(3)
    for i in range(2):
        tp, tm = th.copy(), th.copy()
        tp[i] += eps; tm[i] -= eps
        pp, pm = lml_of(*tp)[1], lml_of(*tm)[1]
        dA = (np.array([pp.expm(x) for x in dt]) - np.array([pm.expm(x) for x in dt])) / (2 * eps)
        dP = (pp.P_inf() - pm.P_inf()) / (2 * eps)
        dQ = (np.array([pp.Q(x, pp.expm(x), pp.P_inf()) for x in dt])
              - np.array([pm.Q(x, pm.expm(x), pm.P_inf()) for x in dt])) / (2 * eps)
        grad[i] = np.sum(g["gA"] * dA) + np.sum(g["gQ"] * dQ) + np.sum(g["gP0"] * dP)
    grad[2] = np.sum(g["gR"])
    for i in range(3):
        tp, tm = th.copy(), th.copy()
        h = 1e-5
        tp[i] += h; tm[i] -= h
        fd = (lml_of(*tp)[0] - lml_of(*tm)[0]) / (2 * h)
        assert abs(grad[i] - fd) <= 1e-6 * max(abs(fd), 1.0), (i, grad[i], fd)


def test_matern_chain_matches_autograd():
    """(gA, gQ) -> (glam, gPinf) for A = blockdiag(expm(F(lam_b) dt)), Q = Pinf - A Pinf A^T."""
    rng = np.random.default_rng(7)
    T, nblk, s = 10, 2, 4
    d = nblk * s
    lam = np.array([1.3, 2.1])
    dt = np.hstack([0.0, rng.uniform(0.05, 0.3, T - 1)])

    def F_of(l):
        F = np.diag(np.ones(s - 1), 1)
        F[-1] = [-l ** 4, -4 * l ** 3, -6 * l ** 2, -4 * l]
        return F

    def dF_of(l):
        D = np.zeros((s, s))
        D[-1] = [-4 * l ** 3, -12 * l ** 2, -12 * l, -4.0]
        return D

    blocks = lambda x: sla.block_diag(*[sla.expm(F_of(l) * x) for l in lam])      # noqa: E731
    dA = lambda b, x: (sla.expm_frechet(F_of(lam[b]) * x, dF_of(lam[b]) * x)[1] if x > 0      # noqa: E731
                       else np.zeros((s, s)))
    # the stationary covariance of the two Matern-7/2 blocks (any other Pinf makes Q_k indefinite)
    Pinf = sla.block_diag(*[sde.Matern72(np.sqrt(7.0) / l, v).to_ss()[5] for l, v in zip(lam, (0.9, 0.4))])
    H = rng.normal(size=(1, d))
    R = np.tile(np.array([[0.3]]), [T, 1, 1])
    Y = rng.normal(size=(T, 1))
    Y[4] = np.nan
    m0 = np.zeros((d, 1))
    A = np.array([blocks(x) for x in dt])
    Q = np.array([Pinf - a @ Pinf @ a.T for a in A])
    g = adjoint.filter_lml_vjp(A, Q, H, R, Y, m0, Pinf, jitter=1e-5)
    glam, gPinf = adjoint.matern_chain(blocks, lam, dt, Pinf, g["gA"], g["gQ"], dA)
    gPinf = gPinf + g["gP0"]                                   # P0 = Pinf as well

    lt = torch.tensor(lam, dtype=torch.float64, requires_grad=True)
    Pt = torch.tensor(Pinf, dtype=torch.float64, requires_grad=True)

    def tF(l):
        top = torch.tensor(np.diag(np.ones(s - 1), 1)[:-1], dtype=torch.float64)
        last = torch.stack([-l ** 4, -4 * l ** 3, -6 * l ** 2, -4 * l])[None]
        return torch.cat([top, last], 0)

    At = torch.stack([torch.block_diag(*[torch.linalg.matrix_exp(tF(lt[b]) * float(x)) for b in range(nblk)])
                      for x in dt])
    Qt = torch.stack([Pt - a @ Pt @ a.T for a in At])
    lml = _torch_filter_lml(At, Qt, torch.tensor(H), torch.tensor(R), Y, torch.tensor(m0), Pt, 1e-5)
    lml.backward()
    assert np.abs(glam - lt.grad.numpy()).max() <= 1e-9 * np.abs(lt.grad.numpy()).max()
    assert np.abs(gPinf - Pt.grad.numpy()).max() <= 1e-9 * np.abs(Pt.grad.numpy()).max()
