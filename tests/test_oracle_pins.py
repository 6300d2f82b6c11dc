"""Derived known-answer pins for the oracle: the reference ships no tests,
so these are the checks its own code supports by construction."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import dense_gp, filters, sde

KERNELS = [
    ("m32", lambda: sde.Matern32(1.0, 1.3)),
    ("m52", lambda: sde.Matern52(0.8, 0.7)),
    ("m72", lambda: sde.Matern72(1.2, 1.1)),
    ("sum", lambda: sde.SumKernel([sde.Matern72(1.2, 1.0), sde.Matern32(0.5, 0.5)])),
]


@pytest.mark.parametrize("name,mk", KERNELS)
def test_expm_closed_form_matches_scipy(name, mk):
    k = mk()
    F = k.to_ss()[0]
    for dt in (0.0, 0.05, 0.37, 0.9):
        np.testing.assert_allclose(k.expm(dt), sla.expm(F * dt), rtol=0, atol=2e-14)


@pytest.mark.parametrize("name,mk", KERNELS)
def test_pinf_solves_lyapunov(name, mk):
    F, L, Qc, H, minf, Pinf = mk().to_ss()
    res = sde.lyapunov_residual(F, L, Qc, Pinf)
    assert np.abs(res).max() <= 1e-14 * np.abs(F @ Pinf).max()


@pytest.mark.parametrize("name,mk", KERNELS)
def test_filter_lml_and_smoother_match_dense_gp(name, mk):
    """With jitter = 0 the sequential filter's lml is log N(y | 0, K + s2 I) and the smoother is the
    dense GP posterior (log_marginal_likelihoods.py:36-58)."""
    rng = np.random.default_rng(1)
    T = 150
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    y = rng.normal(size=T)
    k = mk()
    prior = sde.LTI_SDE([k])
    R = np.tile(np.array([[0.1]]), [T, 1, 1])
    lml, mf, Pf, _ = filters.filter_sequential(prior, t, y[:, None], R, jitter=0.0)
    ms, Ps = filters.smoother_sequential(prior, t, mf, Pf, jitter=0.0)
    Kd = k.K(t, t)
    assert abs(lml - dense_gp.log_marginal_likelihood(Kd, y, 0.1)) <= 1e-10 * abs(lml)
    pm, pv = dense_gp.posterior(Kd, y, 0.1)
    np.testing.assert_allclose(ms[:, 0, 0], pm, rtol=0, atol=1e-11)
    np.testing.assert_allclose(Ps[:, 0, 0], pv, rtol=0, atol=1e-11)


def test_missing_data_equals_dropping_points():
    """NaN-masked steps must give the lml of the observed subset (gaussian.py:72-108)."""
    rng = np.random.default_rng(2)
    T = 120
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    y = rng.normal(size=T)
    miss = rng.uniform(size=T) < 0.2
    y_nan = y.copy()
    y_nan[miss] = np.nan
    k = sde.Matern52(0.8, 0.7)
    prior = sde.LTI_SDE([k])
    R = np.tile(np.array([[0.2]]), [T, 1, 1])
    lml, _, _, _ = filters.filter_sequential(prior, t, y_nan[:, None], R, jitter=0.0)
    Kd = k.K(t[~miss], t[~miss])
    assert abs(lml - dense_gp.log_marginal_likelihood(Kd, y[~miss], 0.2)) <= 1e-10 * abs(lml)


def test_parallel_reference_quirk_q1_is_the_only_difference():
    """At jitter = 0, filter('parallel') differs from the sequential filter only through element 0's
    doubled prior: a sequential filter started from 2 Pinf reproduces it.  (With
    jitter > 0 the two also place the jitter differently: (I - KH)Q vs Q - K S K^T.)"""
    rng = np.random.default_rng(3)
    T = 60
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    y = rng.normal(size=(T, 1))
    k = sde.Matern32(1.0, 1.3)
    prior = sde.LTI_SDE([k])
    R = np.tile(np.array([[0.1]]), [T, 1, 1])
    _, m_par, P_par = filters.filter_parallel_reference(prior, t, y, R, jitter=0.0)

    # sequential filter started from 2 Pinf (A_0 = I, Q_0 := Pinf)
    Pinf = prior.P_inf()
    m, P = prior.m_inf(), 2 * Pinf
    H = prior.H()
    ms, Ps = [], []
    dt = np.hstack([0, np.diff(t)])
    for kk in range(T):
        if kk > 0:
            A = prior.expm(dt[kk])
            m, P = A @ m, A @ P @ A.T + prior.Q(dt[kk], A, Pinf)
        m, P, _ = filters.kf_update_step(m, P, H, R[kk], y[kk][:, None], jitter=0.0)
        ms.append(m), Ps.append(P)
    np.testing.assert_allclose(m_par, np.array(ms), rtol=0, atol=1e-9)
    np.testing.assert_allclose(P_par, np.array(Ps), rtol=0, atol=1e-9)
