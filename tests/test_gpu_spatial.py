"""-m gpu: the spatial conditional after the smoother (`physs_spatial_conditional_f64`) through the host
mirror `physs_gp_b200.spatial` against vectors produced by the reference's own `gaussian_spatial_conditional_cholesky`
(tests/golden/make_golden_spatial.py), against the numpy oracle at config-2 size, and end to end behind the smoother."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "spatial_cond_*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_spatial_conditional_matches_reference_vectors(cuda_device, path):
    from physs_gp_b200 import spatial
    g = np.load(path)
    jit = float(g["jitter"])
    Ktt = g["Ktt"] if np.ptp(g["Ktt"]) > 0 else float(g["Ktt"][0])
    mu, var = spatial.spatial_conditional_block(g["Kzz"], g["Ksz"], g["Kss"], Ktt, g["pred_mean"], g["pred_var"],
                                                jitter=jit)
    assert tuple(mu.shape) == g["mu"].shape and tuple(var.shape) == g["var"].shape
    assert rel(mu, g["mu"]) < TOL and rel(var, g["var"]) < TOL
    v = var.cpu().numpy()
    assert np.abs(v - np.swapaxes(v, -1, -2)).max() < 1e-13 * np.abs(v).max()    # mirrored lower-triangle product
    mu_d, var_d = spatial.spatial_conditional_block(g["Kzz"], g["Ksz"], g["Kss"], Ktt, g["pred_mean"], g["pred_var"],
                                                    diagonal=True, jitter=jit)
    assert rel(mu_d, g["mu"]) < TOL
    assert rel(var_d[..., 0], np.diagonal(g["var"][:, 0], axis1=-2, axis2=-1)) < TOL


def _gram(X1, X2, ls, var):
    r = np.sqrt(((X1[:, None, :] - X2[None, :, :]) ** 2).sum(-1))
    a = np.sqrt(3.0) * r / ls
    return var * (1.0 + a) * np.exp(-a)


@pytest.mark.parametrize("M,N,T", [(200, 200, 300), (200, 333, 5), (37, 1, 3), (1, 5, 2), (208, 64, 150)])
def test_cuda_spatial_conditional_matches_oracle(cuda_device, M, N, T):
    """config-2 size (M = 200 inducing points), more steps than SMs (persistent loop), ragged tile edges."""
    from oracle import dense_gp
    from physs_gp_b200 import spatial
    rng = np.random.default_rng(M * 1000 + N)
    X, XS = rng.uniform(size=[M, 2]), rng.uniform(size=[N, 2])
    Kzz, Ksz, Kss = _gram(X, X, 0.3, 1.1), _gram(XS, X, 0.3, 1.1), _gram(XS, XS, 0.3, 1.1)
    pm = rng.normal(size=[T, M, 1])
    B = rng.normal(size=[T, M, M]) * 0.1
    pv = 0.2 * Kzz[None] + B @ np.swapaxes(B, 1, 2) / M + 0.01 * np.eye(M)
    Ktt = 0.8 + 0.2 * rng.uniform(size=T)
    jit = 1e-6
    mu_o, var_o = dense_gp.spatial_conditional(Kzz, Ksz, Kss, Ktt, pm, pv, jit)
    mu, var = spatial.spatial_conditional_block(Kzz, Ksz, Kss, Ktt, pm, pv, jitter=jit)
    assert rel(mu, mu_o) < TOL and rel(var, var_o) < TOL
    _, var_d = spatial.spatial_conditional_block(Kzz, Ksz, Kss, Ktt, pm, pv, diagonal=True, jitter=jit)
    assert rel(var_d[..., 0], np.diagonal(var_o[:, 0], axis1=-2, axis2=-1)) < TOL


def test_spatial_prediction_behind_the_smoother(cuda_device, monkeypatch):
    """End to end: separable prior -> filter + smoother (f only) -> spatial conditional at new points, against the
    numpy oracle's filter / smoother + dense conditional; at the training points themselves the conditional returns
    the smoothed marginals."""
    from oracle import dense_gp, filters as ofilters, sde as osde
    from physs_gp_b200 import data, kernels as K, likelihood, models, sdes, settings, spatial
    monkeypatch.setattr(settings, "jitter", 1e-6)
    rng = np.random.default_rng(7)
    Ns, T = 24, 40
    X = rng.uniform(size=[Ns, 2])
    XS = np.vstack([X[:5], rng.uniform(size=[9, 2])])
    kern = lambda A, B: _gram(A, B, 0.4, 1.0)
    Ks = kern(X, X)
    t = np.cumsum(rng.uniform(0.05, 0.15, T))
    Y = rng.normal(size=[T, 1, Ns])
    Y[rng.uniform(size=Y.shape) < 0.1] = np.nan
    prior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(K.Matern32(0.6, 0.9), Ks)]))
    model = models.SDE_GP(data.TemporalData(t, Y, X), prior, likelihood.Gaussian(0.05))
    mu, var = spatial.spatial_conditional(model, XS, kern, diagonal=False)
    oprior = osde.LTI_SDE([osde.SpaceTimeSeparable(osde.Matern32(0.6, 0.9), Ks)])
    R = np.tile(0.05 * np.eye(Ns), [T, 1, 1])
    _, mf, Pf, _ = ofilters.filter_sequential(oprior, t, Y[:, 0, :], R, 1e-6)
    ms, Ps = ofilters.smoother_sequential(oprior, t, mf, Pf, full_state=False, jitter=1e-6)
    mu_o, var_o = dense_gp.spatial_conditional(Ks, kern(XS, X), kern(XS, XS), np.full(T, 0.9), ms, Ps, 1e-6)
    assert rel(mu, mu_o) < 1e-8 and rel(var, var_o) < 1e-8
    assert np.abs(mu.cpu().numpy()[:, :5, 0] - ms[:, :5, 0]).max() < 1e-4     # jittered Kzz: not exact


@pytest.mark.parametrize("Nz,Nx", [(12, 20), (40, 33)])
def test_spatial_sparsity_cvi_iterations_match_oracle(cuda_device, monkeypatch, Nz, Nx):
    """SpatialSparsity CVI (sites at Nz inducing points, data at Nx other points, Gaussian likelihood with missing
    entries): two natural-gradient iterations and the ELBO against the numpy oracle -- filter / smoother on the
    separable prior, ELL gradients pulled back through the spatial conditional, block update.  Nz = 12 runs the
    small-block site kernel, Nz = 40 the large-block one."""
    from oracle import cvi as ocvi, filters as ofilters, sde as osde
    from physs_gp_b200 import cvi, kernels as K, sdes, settings, spatial
    monkeypatch.setattr(settings, "jitter", 1e-5)
    rng = np.random.default_rng(100 + Nz)
    T, beta, s2 = 11, 0.6, 0.2
    Z, X = rng.uniform(size=[Nz, 2]), rng.uniform(size=[Nx, 2])
    Kzz, Kxz, Kxx = _gram(Z, Z, 0.4, 1.0), _gram(X, Z, 0.4, 1.0), _gram(X, X, 0.4, 1.0)
    t = np.cumsum(rng.uniform(0.05, 0.15, T))
    Y = rng.normal(size=[T, Nx])
    Y[rng.uniform(size=Y.shape) < 0.15] = np.nan
    kvar = 0.9
    pprior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(K.Matern32(0.7, kvar), Kzz)]))
    oprior = osde.LTI_SDE([osde.SpaceTimeSeparable(osde.Matern32(0.7, kvar), Kzz)])
    q = cvi.FullConjugateGaussian(t, pprior, Nz, B=1)
    model = spatial.SpatialSparsityVGP(Y, s2, q, Kzz, Kxz, Kxx, Ktt=kvar)
    for _ in range(2):
        model.natural_gradient_update(beta)
    elbo = model.elbo()
    torch.cuda.synchronize()
    W, C0 = spatial.conditional_weights(Kzz, Kxz, Kxx, 1e-5)
    c0 = kvar * np.diag(C0)
    Yt = 1e-5 * np.ones((T, Nz)); Vt = np.tile(np.eye(Nz), [T, 1, 1])

    def grads(qm, qv):
        out = [ocvi.spatial_sparsity_gaussian_ell_and_grads(Y[i], s2, W, c0, 1e-5, qm[i][:, 0], qv[i]) for i in range(T)]
        return sum(o[0] for o in out), np.stack([o[1] for o in out]), np.stack([o[2] for o in out])
    for _ in range(2):
        _, qm, qv = ofilters.filter_and_smooth(oprior, t, Yt, Vt)
        _, dm, dS = grads(qm, qv)
        Yt, Vt = ocvi.cvi_step(Yt, Vt, qm[:, :, 0], qv, dm, dS, beta)
    lml, qm, qv = ofilters.filter_and_smooth(oprior, t, Yt, Vt)
    ref = ocvi.elbo(grads(qm, qv)[0], ocvi.surrogate_ell(Yt, Vt, qm[:, :, 0], qv), lml)
    assert rel(model.q.Y_tilde[0], Yt) < 1e-8 and rel(model.q.V_tilde[0], Vt) < 1e-8
    assert abs(float(elbo[0]) - ref) <= 1e-8 * abs(ref)


@pytest.mark.parametrize("kind,Nz,Nx", [("poisson", 12, 20), ("bernoulli", 10, 17), ("poisson", 40, 33)])
def test_spatial_sparsity_non_gaussian_cvi_matches_oracle(cuda_device, monkeypatch, kind, Nz, Nx):
    """SpatialSparsity CVI with independent Poisson / Bernoulli observations AWAY from the inducing points: the
    per-point Gauss-Hermite terms come from the site kernel on N scalar blocks per step, are pulled back through the
    spatial conditional (W^T E[l'], W^T diag(1/2 E[l'']) W) and feed the ordinary block update; two iterations and the
    ELBO against the numpy oracle (whose pull-back is checked against central differences in tests/test_oracle_cvi.py)."""
    from oracle import cvi as ocvi, filters as ofilters, sde as osde
    from physs_gp_b200 import cvi, kernels as K, sdes, settings, spatial
    monkeypatch.setattr(settings, "jitter", 1e-5)
    rng = np.random.default_rng(300 + Nz)
    T, beta = 9, 0.5
    Z, X = rng.uniform(size=[Nz, 2]), rng.uniform(size=[Nx, 2])
    Kzz, Kxz, Kxx = _gram(Z, Z, 0.4, 1.0), _gram(X, Z, 0.4, 1.0), _gram(X, X, 0.4, 1.0)
    t = np.cumsum(rng.uniform(0.05, 0.15, T))
    Y = (rng.integers(0, 4, size=[T, Nx]) if kind == "poisson" else rng.integers(0, 2, size=[T, Nx])).astype(float)
    Y[rng.uniform(size=Y.shape) < 0.15] = np.nan
    kvar = 0.9
    pprior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(K.Matern32(0.7, kvar), Kzz)]))
    oprior = osde.LTI_SDE([osde.SpaceTimeSeparable(osde.Matern32(0.7, kvar), Kzz)])
    q = cvi.FullConjugateGaussian(t, pprior, Nz, B=1)
    lik = cvi.PoissonLik(1.0) if kind == "poisson" else cvi.BernoulliLik()
    model = spatial.SpatialSparsityVGP(Y, None, q, Kzz, Kxz, Kxx, Ktt=kvar, likelihood=lik)
    for _ in range(2):
        model.natural_gradient_update(beta)
    elbo = model.elbo()
    torch.cuda.synchronize()
    W, C0 = spatial.conditional_weights(Kzz, Kxz, Kxx, 1e-5)
    c0 = kvar * np.diag(C0)
    Yt = 1e-5 * np.ones((T, Nz)); Vt = np.tile(np.eye(Nz), [T, 1, 1])

    def grads(qm, qv):
        out = [ocvi.spatial_sparsity_gh_ell_and_grads(Y[i], kind, W, c0, 1e-5, qm[i][:, 0], qv[i]) for i in range(T)]
        return sum(o[0] for o in out), np.stack([o[1] for o in out]), np.stack([o[2] for o in out])
    for _ in range(2):
        _, qm, qv = ofilters.filter_and_smooth(oprior, t, Yt, Vt)
        _, dm, dS = grads(qm, qv)
        Yt, Vt = ocvi.cvi_step(Yt, Vt, qm[:, :, 0], qv, dm, dS, beta)
    lml, qm, qv = ofilters.filter_and_smooth(oprior, t, Yt, Vt)
    ref = ocvi.elbo(grads(qm, qv)[0], ocvi.surrogate_ell(Yt, Vt, qm[:, :, 0], qv), lml)
    assert rel(model.q.Y_tilde[0], Yt) < 1e-7 and rel(model.q.V_tilde[0], Vt) < 1e-7
    assert abs(float(elbo[0]) - ref) <= 1e-8 * abs(ref)
