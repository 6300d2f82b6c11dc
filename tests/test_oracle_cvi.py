"""Derived pins for the CVI oracle."""
import numpy as np

from oracle import cvi, filters, sde
from tests import synth


def test_theta_lambda_round_trip():
    rng = np.random.default_rng(0)
    D = 4
    V = synth.random_spd(rng, (), D)
    Y = rng.normal(size=(D, 1))
    l1, l2 = cvi.theta_to_lambda(Y, V, ng_jitter=0.0)
    t1, t2 = cvi.lambda_to_theta(l1, l2, ng_jitter=0.0)
    np.testing.assert_allclose(t1, Y, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(t2, V, rtol=1e-10, atol=1e-12)


def test_gaussian_ell_gradients_match_finite_differences():
    rng = np.random.default_rng(1)
    D, P = 4, 2
    W = rng.normal(size=(P, D))
    noise = synth.random_spd(rng, (), P)
    q_mu = rng.normal(size=D)
    q_var = synth.random_spd(rng, (), D)
    for Y in (rng.normal(size=P), np.array([0.3, np.nan])):
        ell, dm, dS = cvi.gaussian_ell_and_grads(Y, noise, W, q_mu, q_var)
        eps = 1e-6
        for i in range(D):
            e = np.zeros(D); e[i] = eps
            fd = (cvi.gaussian_ell_and_grads(Y, noise, W, q_mu + e, q_var)[0]
                  - cvi.gaussian_ell_and_grads(Y, noise, W, q_mu - e, q_var)[0]) / (2 * eps)
            assert abs(fd - dm[i]) < 1e-6 * max(1.0, abs(dm[i]))
        for i in range(D):
            for j in range(D):
                E = np.zeros((D, D)); E[i, j] = eps
                fd = (cvi.gaussian_ell_and_grads(Y, noise, W, q_mu, q_var + E)[0]
                      - cvi.gaussian_ell_and_grads(Y, noise, W, q_mu, q_var - E)[0]) / (2 * eps)
                assert abs(fd - dS[i, j]) < 1e-6 * max(1.0, abs(dS[i, j]))


def test_gh_matches_poisson_closed_form_and_fd():
    rng = np.random.default_rng(2)
    for _ in range(5):
        y = float(rng.integers(0, 8)); m = rng.normal() * 0.5; v = rng.uniform(0.05, 0.6)
        ell, d1, d2 = cvi.gh_ell_and_grads(y, m, v, "poisson", K=20, binsize=0.7)
        cf = cvi.poisson_ell_closed_form(y, m, v, 0.7)
        assert abs(ell - cf) < 1e-10 * max(1.0, abs(cf))
        eps = 1e-5
        fd_m = (cvi.poisson_ell_closed_form(y, m + eps, v, 0.7) - cvi.poisson_ell_closed_form(y, m - eps, v, 0.7)) / (2 * eps)
        fd_v = (cvi.poisson_ell_closed_form(y, m, v + eps, 0.7) - cvi.poisson_ell_closed_form(y, m, v - eps, 0.7)) / (2 * eps)
        assert abs(d1 - fd_m) < 1e-7 and abs(d2 - fd_v) < 1e-7


def test_gh_bernoulli_converges_with_order():
    y, m, v = 1.0, 0.3, 0.5
    a = cvi.gh_ell_and_grads(y, m, v, "bernoulli", K=20)
    b = cvi.gh_ell_and_grads(y, m, v, "bernoulli", K=64)
    assert max(abs(x - z) for x, z in zip(a, b)) < 1e-6


def test_cvi_gaussian_fixed_point_and_elbo_equals_lml():
    """With a Gaussian likelihood and beta = 1 one CVI step returns (Y~, V~) = (y, R) up to
    O(ng_jitter), and the CVI ELBO then equals the exact marginal likelihood to the same order
    (cvi_nat_grad.py:62-73, elbos.py:163-194)."""
    rng = np.random.default_rng(3)
    T = 60
    k = sde.Matern32(1.0, 1.2)
    prior = sde.LTI_SDE([k])                      # sites over f only: D = m = 1
    t = synth.time_grid(T, 0.1, rng)
    y = rng.normal(size=(T, 1))
    Rn = 0.3
    Ytil = 1e-5 * np.ones((T, 1))                 # reference init (conjugate_gaussian_approximate_posterior.py:209-218)
    Vtil = np.tile(np.eye(1), [T, 1, 1])
    _, q_mu, q_var = filters.filter_and_smooth(prior, t, Ytil, Vtil)
    dm = np.empty((T, 1)); dS = np.empty((T, 1, 1))
    for i in range(T):
        _, dm[i], dS[i] = cvi.gaussian_ell_and_grads(y[i], np.array([[Rn]]), None, q_mu[i][:, 0], q_var[i])
    Yn, Vn = cvi.cvi_step(Ytil, Vtil, q_mu[:, :, 0], q_var, dm, dS, beta=1.0, ng_jitter=1e-9)
    np.testing.assert_allclose(Yn, y, rtol=0, atol=1e-7)
    np.testing.assert_allclose(Vn[:, 0, 0], Rn, rtol=0, atol=1e-7)
    # ELBO at the fixed point
    lml_s, q_mu2, q_var2 = filters.filter_and_smooth(prior, t, Yn, Vn, jitter=0.0)
    ell = sum(cvi.gaussian_ell_and_grads(y[i], np.array([[Rn]]), None, q_mu2[i][:, 0], q_var2[i])[0] for i in range(T))
    ell_s = cvi.surrogate_ell(Yn, Vn, q_mu2[:, :, 0], q_var2)
    exact, _, _, _ = filters.filter_sequential(prior, t, y, np.tile(np.array([[Rn]]), [T, 1, 1]), jitter=0.0)
    assert abs(cvi.elbo(ell, ell_s, lml_s) - exact) < 1e-5 * abs(exact)


def test_pendulum_collocation_ell_closed_form_vs_quadrature_and_finite_differences():
    """The closed-form damped-oscillator collocation ELL (oracle.cvi.pendulum_ell_and_grads): value against a
    40^3-point tensor Gauss-Hermite rule, gradients against central finite differences of the closed form."""
    from oracle import cvi as ocvi
    rng = np.random.default_rng(3)
    for trial in range(4):
        G = rng.normal(size=(3, 3)) * 0.3
        S = G @ G.T + 0.02 * np.eye(3)
        m = rng.normal(size=3) * np.array([1.0, 0.5, 0.5])
        y = np.array([m[0] + 0.1, 0.0]) if trial < 3 else np.array([np.nan, 0.0])
        args = (1.3, 0.2, 0.05, 0.01)
        ell, dm, dS = ocvi.pendulum_ell_and_grads(y, m, S, *args)
        assert abs(ell - ocvi.pendulum_ell_quadrature(y, m, S, *args)) < 1e-9 * max(1.0, abs(ell))
        h = 1e-6
        for i in range(3):
            e = np.zeros(3); e[i] = h
            fd = (ocvi.pendulum_ell_and_grads(y, m + e, S, *args)[0] - ocvi.pendulum_ell_and_grads(y, m - e, S, *args)[0]) / (2 * h)
            assert abs(fd - dm[i]) < 1e-6 * max(1.0, abs(dm[i]))
            for j in range(3):
                E = np.zeros((3, 3)); E[i, j] += h / 2; E[j, i] += h / 2      # symmetric perturbation
                fd = (ocvi.pendulum_ell_and_grads(y, m, S + E, *args)[0] - ocvi.pendulum_ell_and_grads(y, m, S - E, *args)[0]) / (2 * h)
                sym = 0.5 * (dS[i, j] + dS[j, i]) if i != j else dS[i, j]
                assert abs(fd - (sym if i == j else 2 * sym * 0.5)) < 1e-6 * max(1.0, abs(sym)), (i, j, fd, sym)
        # Gauss-Newton curvature is negative semi-definite and equals -1/2 J^T J / var
        _, _, dS_gn = ocvi.pendulum_ell_and_grads(y, m, S, *args, gauss_newton=True)
        assert np.linalg.eigvalsh(dS_gn).max() <= 1e-12
        J = np.array([1.3 * np.cos(m[0]), 0.2, 1.0])
        ref = -0.5 * np.outer(J, J) / 0.01 + (0 if np.isnan(y[0]) else -0.5 * np.outer([1, 0, 0], [1, 0, 0]) / 0.05)
        assert np.allclose(dS_gn, ref, rtol=1e-13, atol=1e-13)


def test_vectorised_scalar_cvi_iteration_matches_block_loops():
    """oracle/cvi_vec.py (the CPU restatement timed by bench.py for config 4) == the per-block loops of
    oracle/cvi.py composed as natural_gradients + elbo, on a small Poisson and Bernoulli problem."""
    from oracle import cvi_vec
    rng = np.random.default_rng(11)
    B, T = 3, 40
    t = np.cumsum(rng.uniform(0.5, 1.5, T) * 0.1)
    ls = rng.uniform(0.5, 2.0, B)
    for kind in ("poisson", "bernoulli"):
        Y = (rng.poisson(1.5, size=(B, T)) if kind == "poisson" else rng.integers(0, 2, size=(B, T))).astype(float)
        Y[rng.uniform(size=Y.shape) < 0.1] = np.nan
        Ytil = rng.normal(size=(B, T)) * 0.3
        Vtil = rng.uniform(0.5, 2.0, size=(B, T))
        kerns = [sde.Matern32(l, 1.0) for l in ls]
        lam = np.array([[np.sqrt(3.0) / l] for l in ls])
        Pinf = np.stack([sde.LTI_SDE([k]).P_inf() for k in kerns])
        H = sde.LTI_SDE([kerns[0]]).H()
        Yn, Vn, el = cvi_vec.cvi_iteration((2, lam, Pinf, H), t, Y, Ytil, Vtil, kind, 0.3, K=20, binsize=0.8)
        for b in range(B):
            prior = sde.LTI_SDE([kerns[b]])

            def posterior(Yt, Vt):
                lml, mf, Pf, _ = filters.filter_sequential(prior, t, Yt[:, None], Vt[:, None, None], 1e-5)
                ms, Ps = filters.smoother_sequential(prior, t, mf, Pf, full_state=False, jitter=1e-5)
                return lml, ms[:, :, 0], Ps
            _, qm, qv = posterior(Ytil[b], Vtil[b])
            g = [cvi.gh_ell_and_grads(Y[b, k], qm[k, 0], qv[k, 0, 0], kind, 20, 0.8) for k in range(T)]
            dm = np.array([x[1] for x in g])[:, None]
            dS = np.array([x[2] for x in g])[:, None, None]
            Y1, V1 = cvi.cvi_step(Ytil[b][:, None], Vtil[b][:, None, None], qm, qv, dm, dS, 0.3)
            np.testing.assert_allclose(Yn[b], Y1[:, 0], rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(Vn[b], V1[:, 0, 0], rtol=1e-10, atol=1e-12)
            lml, qm2, qv2 = posterior(Y1[:, 0], V1[:, 0, 0])
            ell = sum(cvi.gh_ell_and_grads(Y[b, k], qm2[k, 0], qv2[k, 0, 0], kind, 20, 0.8)[0] for k in range(T))
            ref = cvi.elbo(ell, cvi.surrogate_ell(Y1, V1, qm2, qv2), lml)
            assert abs(el[b] - ref) < 1e-9 * abs(ref)


def test_gauss_hermite_ell_within_4_sigma_of_reference_monte_carlo():
    """Statistical pin of the quadrature that replaces the reference's Monte-Carlo ELL: K = 20
    Gauss-Hermite vs 2e5 reparameterised samples drawn by the reference's own mv_indepentdent_monte_carlo
    (integrals/approximators.py:16-58) with its own Poisson / Bernoulli log-likelihoods
    (tests/golden/make_golden_cvi.py -> tests/golden/mc_ell.npz), within 4 standard errors."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_ell.npz"))
    for kind in ("poisson", "bernoulli"):
        for y, m, v, mean, se in zip(g[kind + "_y"], g[kind + "_m"], g[kind + "_v"], g[kind + "_mc_mean"],
                                     g[kind + "_mc_stderr"]):
            ell = cvi.gh_ell_and_grads(float(y), float(m), float(v), kind, K=20, binsize=float(g[kind + "_binsize"]))[0]
            assert abs(ell - mean) < 4.0 * se, (kind, y, m, v, ell, mean, se)
            assert se < 0.02 * max(1.0, abs(mean))             # the pin is tight enough to mean something


def test_gauss_hermite_20_equals_poisson_closed_form_for_moderate_variance():
    """Pin behind the one-exp fast path of the Poisson site kernel (csrc/physs_cvi_core.cuh): for sd = sqrt(2 v) <= 3
    the K = 20 Gauss-Hermite sums of l, l', l'' equal the reference's closed-form Poisson ELL
    (expected_log_likelihoods.py:149-174) to round-off, so switching between the two is invisible at 1e-9."""
    x, w = np.polynomial.hermite.hermgauss(20)
    w = w / np.sqrt(np.pi)
    assert np.array_equal(x, -x[::-1])                     # exact +- pairs (the kernel's reciprocal shortcut)
    rng = np.random.default_rng(0)
    for sd in np.concatenate([np.linspace(0.0, 3.0, 31), rng.uniform(0, 3, 50)]):
        m, y, b = rng.normal(), float(rng.integers(0, 6)), 0.7
        f = m + sd * x
        lam = b * np.exp(f)
        E = b * np.exp(m + 0.25 * sd * sd)
        for quad, closed in (((w * (y * f - lam)).sum(), y * m - E), ((w * (y - lam)).sum(), y - E),
                             ((w * -lam).sum(), -E)):
            assert abs(quad - closed) <= 2e-15 * max(1.0, abs(closed), E)


def test_spatial_sparsity_ell_gradients_by_finite_differences():
    """oracle.cvi.spatial_sparsity_gaussian_ell_and_grads: the closed-form pull-back of the data-location gradients
    through the spatial conditional equals central differences of the ELL w.r.t. every entry of (q_mu, q_var) --
    the quantity `jax.grad(partial_ell, (1, 2))` returns (cvi_nat_grad.py:381-383)."""
    import numpy as np
    from oracle import cvi as ocvi
    rng = np.random.default_rng(3)
    M, N = 5, 8
    W = rng.normal(size=(N, M))
    c0 = rng.uniform(0.1, 0.5, N)
    y = rng.normal(size=N)
    y[[1, 6]] = np.nan
    qm = rng.normal(size=M)
    B = rng.normal(size=(M, M))
    qS = B @ B.T + np.eye(M)
    ell, dm, dS = ocvi.spatial_sparsity_gaussian_ell_and_grads(y, 0.3, W, c0, 1e-5, qm, qS)
    h = 1e-6
    for i in range(M):
        e = np.zeros(M); e[i] = h
        fd = (ocvi.spatial_sparsity_gaussian_ell_and_grads(y, 0.3, W, c0, 1e-5, qm + e, qS)[0]
              - ocvi.spatial_sparsity_gaussian_ell_and_grads(y, 0.3, W, c0, 1e-5, qm - e, qS)[0]) / (2 * h)
        assert abs(fd - dm[i]) < 1e-6 * max(1.0, abs(dm[i]))
        for j in range(M):
            E = np.zeros((M, M)); E[i, j] = h
            fd = (ocvi.spatial_sparsity_gaussian_ell_and_grads(y, 0.3, W, c0, 1e-5, qm, qS + E)[0]
                  - ocvi.spatial_sparsity_gaussian_ell_and_grads(y, 0.3, W, c0, 1e-5, qm, qS - E)[0]) / (2 * h)
            assert abs(fd - dS[i, j]) < 1e-6 * max(1.0, abs(dS[i, j]))
    # with the data AT the inducing points (W = I, c0 = 0, no jitter) it reduces to the NoSparsity gradients
    ell1, dm1, dS1 = ocvi.spatial_sparsity_gaussian_ell_and_grads(y[:M], 0.3, np.eye(M), np.zeros(M), 0.0, qm, qS)
    ell0, dm0, dS0 = ocvi.gaussian_ell_and_grads(y[:M], 0.3 * np.eye(M), None, qm, qS)
    assert abs(ell1 - ell0) < 1e-12 * abs(ell0) and np.allclose(dm1, dm0, atol=1e-13) and np.allclose(dS1, dS0, atol=1e-13)


def test_spatial_sparsity_non_gaussian_ell_gradients_by_finite_differences():
    """oracle.cvi.spatial_sparsity_gh_ell_and_grads (Poisson / Bernoulli observations away from the inducing points):
    W^T E[l'] and W^T diag(1/2 E[l'']) W equal central differences of a 60-point quadrature of the ELL w.r.t. every
    entry of (q_mu, q_var) -- what `jax.grad(partial_ell, (1, 2))` returns (cvi_nat_grad.py:381-383).  (Bonnet / Price:
    dE[l]/dm = E[l'], dE[l]/dv = 1/2 E[l'']; a high-order rule makes the quadrature error of both sides negligible.)"""
    import numpy as np
    from oracle import cvi as ocvi
    rng = np.random.default_rng(4)
    M, N, K = 4, 7, 60
    W = 0.5 * rng.normal(size=(N, M))
    c0 = rng.uniform(0.05, 0.2, N)
    qm = 0.3 * rng.normal(size=M)
    B = 0.3 * rng.normal(size=(M, M))
    qS = B @ B.T + 0.2 * np.eye(M)
    for kind in ("poisson", "bernoulli"):
        y = (rng.integers(0, 4, N) if kind == "poisson" else rng.integers(0, 2, N)).astype(float)
        y[2] = np.nan
        f = lambda m, S: ocvi.spatial_sparsity_gh_ell_and_grads(y, kind, W, c0, 1e-5, m, S, K=K)[0]     # noqa: E731
        ell, dm, dS = ocvi.spatial_sparsity_gh_ell_and_grads(y, kind, W, c0, 1e-5, qm, qS, K=K)
        h = 1e-5
        for i in range(M):
            e = np.zeros(M); e[i] = h
            fd = (f(qm + e, qS) - f(qm - e, qS)) / (2 * h)
            assert abs(fd - dm[i]) < 1e-6 * max(1.0, abs(dm[i])), (kind, i)
            for j in range(M):
                E = np.zeros((M, M)); E[i, j] = h
                fd = (f(qm, qS + E) - f(qm, qS - E)) / (2 * h)
                assert abs(fd - dS[i, j]) < 1e-6 * max(1.0, abs(dS[i, j])), (kind, i, j)
