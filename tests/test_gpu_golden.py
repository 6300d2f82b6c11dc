"""-m gpu: the CUDA path (host API -> C ABI) against tests/golden/ -- outputs of the reference's own
source files (tests/golden/README.md) -- with NO oracle in between.  Tolerance 1e-9 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

from tests.golden.cases import CASES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def product_prior(name):
    from physs_gp_b200 import kernels as K
    from physs_gp_b200 import sdes
    latents, fso = CASES[name][0], CASES[name][1]
    kind = {"m32": K.Matern32, "m52": K.Matern52, "m72": K.Matern72}
    lat = [K.sum_kernels([kind[k](ls, var) for k, ls, var in parts]) for parts in latents]
    ind = sdes.Independent(lat)
    return sdes.LTI_SDE_Full_State_Obs(ind) if fso else sdes.LTI_SDE(ind)


@pytest.mark.parametrize("jit", [1e-5, 0.0])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_filter_smoother_matches_reference_vectors(cuda_device, name, jit, monkeypatch):
    from physs_gp_b200 import data, filters, settings
    monkeypatch.setattr(settings, "jitter", jit)
    g = np.load(os.path.join(GOLD, "filter_%s_jit%s.npz" % (name, "1e-5" if jit else "0")))
    prior = product_prior(name)
    d = data.TemporalData(g["t"], g["Y"][:, :, None])
    lml, kf = filters.filter_loop(d, prior, R=g["R"])
    assert abs(float(lml) - float(g["seq_lml"])) <= TOL * abs(float(g["seq_lml"]))
    assert rel(kf['m'], g["seq_mf"]) < TOL and rel(kf['P'], g["seq_Pf"]) < TOL
    for fs in (False, True):
        mu, var = filters.smoother_loop(d, prior, kf, full_state=fs)
        assert rel(mu, g["seq_ms_full%d" % fs]) < TOL and rel(var, g["seq_Ps_full%d" % fs]) < TOL


@pytest.mark.parametrize("jit", [1e-5, 0.0])
@pytest.mark.parametrize("name", ["m32", "m52", "m72", "indep_m32x2"])
def test_cuda_packed_posterior_call_matches_reference_vectors(cuda_device, name, jit, monkeypatch):
    """The posterior-only call (physs_kf_filter_smooth_packed_f64 behind filters.filter_smooth_fused: packed hand-over,
    no filtered outputs) on a batch of 64 copies of each reference case with state dim <= 4: lml and the projected
    smoothed moments of EVERY copy against the reference's own filter_loop / smoother_loop outputs."""
    from physs_gp_b200 import data, filters, settings
    monkeypatch.setattr(settings, "jitter", jit)
    g = np.load(os.path.join(GOLD, "filter_%s_jit%s.npz" % (name, "1e-5" if jit else "0")))
    prior = product_prior(name)
    B = 64
    Y = np.broadcast_to(g["Y"][None, :, :, None], (B,) + g["Y"].shape + (1,)).copy()
    out = filters.filter_smooth_fused(data.TemporalData(g["t"], Y), prior, R=g["R"])
    assert out is not None, "the packed call must cover these shapes"
    lml, mu, var = out
    assert lml.shape == (B,) and mu.shape[0] == B and var.shape[0] == B
    for b in (0, 31, 32, B - 1):
        assert abs(float(lml[b]) - float(g["seq_lml"])) <= TOL * abs(float(g["seq_lml"]))
        assert rel(mu[b], g["seq_ms_full0"]) < TOL and rel(var[b], g["seq_Ps_full0"]) < TOL
    assert torch.equal(mu[0], mu[B - 1]) and torch.equal(var[0], var[B - 1])


def test_cuda_cvi_blocks_match_reference_vectors(cuda_device):
    """theta -> lambda -> cvi_block_update -> theta (one fused kernel) and the closed-form block ELL."""
    from physs_gp_b200 import cvi
    g = np.load(os.path.join(GOLD, "cvi_blocks.npz"))

    def dev(x):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()
    for D in (1, 3, 6):
        for tag, ngj in (("1e-7", 1e-7), ("1e-5", 1e-5)):
            k = "D%d_ngj%s_" % (D, tag)
            # reference: lambda' = cvi_block_update(theta_to_lambda(Y~, V~), ...); theta' = lambda_to_theta(lambda')
            from oracle import cvi as ocvi   # only the (pinned) lambda -> theta of the reference's n1, n2
            t1, t2 = ocvi.lambda_to_theta(g[k + "n1"], g[k + "n2"], ngj)
            Yn, Vn = cvi.natgrad_step(dev(g[k + "Yt"][None, :, 0]), dev(g[k + "V"][None]), dev(g[k + "mq"][None, :, 0]),
                                      dev(g[k + "S"][None]), None, None, None, float(g[k + "beta"]), ng_jitter=ngj,
                                      dm=dev(g[k + "dm"][None, :, 0]), dS=dev(g[k + "dS"][None]))
            cond = np.linalg.cond(g[k + "V"]) * np.linalg.cond(-2 * g[k + "n2"] + ngj * np.eye(D))
            assert rel(Vn[0], t2) < TOL * max(1.0, cond) and rel(Yn[0], t1[:, 0]) < TOL * max(1.0, cond)
            ell = cvi.expected_log_likelihood(dev(g[k + "mq"][None, :, 0]), dev(g[k + "S"][None]),
                                              dev(g[k + "Yobs"][None, :, 0]), None, cvi.GaussianLik(np.eye(D)),
                                              noise=dev(g[k + "V"][None]))
            assert abs(float(ell[0]) - float(g[k + "ell"])) <= TOL * abs(float(g[k + "ell"]))


def test_cuda_cvi_precision_blocks_match_reference_vectors(cuda_device):
    """'NG_Precision' site update (physs_cvi_natgrad_step_prec_f64; register kernel D <= 4, lane-group kernel above)
    against the reference's own theta_precision_to_lambda -> cvi_block_update -> lambda_to_theta_precision
    (tests/golden/make_golden_prec.py), and the surrogate variance mat_inv(precision)."""
    from physs_gp_b200 import cvi, likelihood

    def dev(x):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()
    g = np.load(os.path.join(GOLD, "cvi_blocks_prec.npz"))
    for D in (1, 2, 3, 4, 6, 8):
        for tag, ngj in (("1e-7", 1e-7), ("1e-5", 1e-5)):
            k = "D%d_ngj%s_" % (D, tag)
            Yn, Pn = cvi.natgrad_step(dev(g[k + "Yt"][None, :, 0]), dev(g[k + "Lam"][None]), dev(g[k + "mq"][None, :, 0]),
                                      dev(g[k + "S"][None]), None, None, None, float(g[k + "beta"]), ng_jitter=ngj,
                                      dm=dev(g[k + "dm"][None, :, 0]), dS=dev(g[k + "dS"][None]), precision=True)
            cond = np.linalg.cond(g[k + "Lam"]) * np.linalg.cond(g[k + "t2"] + ngj * np.eye(D))
            assert rel(Pn[0], g[k + "t2"]) < TOL and rel(Yn[0], g[k + "t1"][:, 0]) < TOL * max(1.0, cond)
            var = likelihood.PrecisionBlockDiagonalGaussian(dev(g[k + "Lam"][None])).variance
            assert rel(var[0], g[k + "var"]) < TOL * max(1.0, np.linalg.cond(g[k + "Lam"]))


def test_cuda_cvi_iteration_and_elbo_match_reference_assembly(cuda_device, monkeypatch):
    """One whole CVI iteration and the ELBO through the reference-shaped objects (cvi.VGP.natural_gradient_update
    / .elbo: posterior kernels -> fused site kernel -> ELL kernels) against the reference's OWN
    `natural_gradients` + `elbo` (tests/golden/make_golden_cvi.py; Gaussian likelihood, full-state sites)."""
    from physs_gp_b200 import cvi, kernels as K, sdes, settings
    g = np.load(os.path.join(GOLD, "cvi_assembly.npz"))
    keys = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_elbo")})
    kind = {"m32": K.Matern32, "m52": K.Matern52, "m72": K.Matern72}
    for key in keys:
        monkeypatch.setattr(settings, "jitter", float(g[key + "_jitter"]))
        monkeypatch.setattr(settings, "ng_jitter", float(g[key + "_ng_jitter"]))
        ls, var = g[key + "_hyper"]
        prior = sdes.LTI_SDE_Full_State_Obs(sdes.Independent([kind[str(g[key + "_kernel"][0])](float(ls), float(var))]))
        Ytil, Vtil = g[key + "_Ytil"], g[key + "_Vtil"]
        T, D = Ytil.shape
        q = cvi.FullConjugateGaussian(g[key + "_t"], prior, D, B=1, Y_tilde=Ytil[None], V_tilde=Vtil[None],
                                      device=cuda_device)
        model = cvi.VGP(g[key + "_Yobs"][None], cvi.GaussianLik(g[key + "_noise"]), q)
        # the ELBO the generator recorded is evaluated at the ORIGINAL sites; evaluate it before the update
        elbo = float(model.elbo()[0])
        assert abs(elbo - float(g[key + "_elbo"])) <= TOL * abs(float(g[key + "_elbo"])), key
        model.natural_gradient_update(float(g[key + "_beta"]))
        cond = max(np.linalg.cond(Vtil[k]) for k in range(T))
        assert rel(q.Y_tilde[0], g[key + "_Ytil_new"]) < TOL * max(1.0, cond), key
        assert rel(q.V_tilde[0], g[key + "_Vtil_new"]) < TOL * max(1.0, cond), key


def test_cuda_gauss_newton_curvature_matches_reference_assembly(cuda_device):
    """The collocation kernel's Gauss-Newton curvature (physs_cvi_ell_pendulum_f64, gauss_newton=1) and the
    generic physs_cvi_gauss_newton_f64 on the recorded Jacobians, against 0.5 * the reference's own assembly lines
    (cvi_hessian_approximations.py slice, see tests/golden/make_golden_cvi.py)."""
    from physs_gp_b200 import cvi
    g = np.load(os.path.join(GOLD, "gn_pendulum.npz"))

    def dev(x):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()
    u, Y, H = g["u"], g["Y"], g["approx_hessian"]
    T, D = u.shape
    lik = cvi.DampedPendulumLik(g=float(g["g"]), l=float(g["l"]), b=float(g["b"]), var_obs=float(g["var_obs"]),
                                var_col=float(g["var_col"]))
    q_var = np.tile(0.1 * np.eye(D), [T, 1, 1])
    _, _, dS = cvi.pendulum_expected_log_likelihood(dev(u), dev(q_var), dev(Y), lik, gauss_newton=True, want_grads=True)
    assert rel(dS, H) < 1e-12
    dS2 = cvi.gauss_newton_curvature(dev(g["J"]), np.array([float(g["var_obs"]), float(g["var_col"])]), y=dev(Y))
    assert rel(dS2, H) < 1e-12


def test_device_sort_pad_matches_reference_vectors(cuda_device):
    """the sort / pad pre-step on the GPU (torch.unique = device sort) against the reference's own
    numpy helpers."""
    from tests.test_golden import _check_sort_pad
    _check_sort_pad(cuda_device)


@pytest.mark.parametrize("q", [1, 2, 3])
def test_cuda_integrated_wiener_matches_reference_vectors(cuda_device, q, monkeypatch):
    """the integrated Wiener prior with A_k, Q_k evaluated ON CHIP (PHYSS_DISC_IWP) against the
    reference's own WienerVelocity under its sequential filter / smoother; the DISC_GIVEN route (A_k, Q_k from the
    host mirror) must agree with it."""
    from physs_gp_b200 import data, filters, kernels as K, ops, sdes, settings, _lib
    g = np.load(os.path.join(GOLD, "iwp_q%d.npz" % q))
    monkeypatch.setattr(settings, "jitter", float(g["jitter"]))
    prior = sdes.LTI_SDE(sdes.Independent([K.WienerVelocity(q, float(g["variance"]), float(g["stable_state_covariance"]))]))
    d = data.TemporalData(g["t"], g["Y"][:, :, None])
    dev = cuda_device
    (disc,), _, _, _ = filters.lower_prior(prior, None, [torch.zeros(3, dtype=torch.float64, device=dev)], dev)
    assert disc.mode == _lib.DISC_IWP                      # the on-chip route is the one under test
    lml, kf = filters.filter_loop(d, prior, R=g["R"])
    assert abs(float(lml) - float(g["lml"])) <= TOL * abs(float(g["lml"]))
    assert rel(kf['m'], g["mf"]) < TOL and rel(kf['P'], g["Pf"]) < TOL
    for fs in (False, True):
        mu, var = filters.smoother_loop(d, prior, kf, full_state=fs)
        assert rel(mu, g["ms_full%d" % fs]) < TOL and rel(var, g["Ps_full%d" % fs]) < TOL
    # parallel-in-time route lowers the same prior to DISC_GIVEN
    lml_p, kf_p = filters.filter_loop(d, prior, R=g["R"], filter_type="b200_parallel")
    assert abs(float(lml_p) - float(g["lml"])) <= TOL * abs(float(g["lml"]))
    assert rel(kf_p['m'], g["mf"]) < TOL and rel(kf_p['P'], g["Pf"]) < TOL


# --------------------------------------------------------------------------- periodic prior (row a6)
import glob  # noqa: E402


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "periodic_*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_periodic_prior_matches_reference_vectors(cuda_device, path, monkeypatch):
    """the periodic prior (stack of harmonic oscillators, optionally summed with a Matern-3/2) with
    A_k = rotation stack evaluated ON CHIP (sign-bit lam blocks of PHYSS_DISC_MATERN) against the reference's own
    ApproxSDEPeriodic_BN -- which goes through the generic Pade expm -- under its sequential filter / smoother;
    the DISC_GIVEN route (A_k, Q_k from the host mirror) and the parallel-in-time kernels must agree with it."""
    from physs_gp_b200 import data, filters, kernels as K, sdes, settings, _lib
    from tests.test_golden import periodic_priors
    g = np.load(path)
    monkeypatch.setattr(settings, "jitter", float(g["jitter"]))
    _, pk = periodic_priors(g, None, K)
    prior = sdes.LTI_SDE(sdes.Independent([pk]))
    d = data.TemporalData(g["t"], g["Y"][:, :, None])
    dev = cuda_device
    (disc,), _, _, _ = filters.lower_prior(prior, None, [torch.zeros(3, dtype=torch.float64, device=dev)], dev)
    assert disc.mode == _lib.DISC_MATERN                   # the on-chip route is the one under test
    lml, kf = filters.filter_loop(d, prior, R=g["R"])
    assert abs(float(lml) - float(g["lml"])) <= TOL * abs(float(g["lml"]))
    assert rel(kf['m'], g["mf"]) < TOL and rel(kf['P'], g["Pf"]) < TOL
    for fs in (False, True):
        mu, var = filters.smoother_loop(d, prior, kf, full_state=fs)
        assert rel(mu, g["ms_full%d" % fs]) < 1e-8 and rel(var, g["Ps_full%d" % fs]) < 1e-8
    lml_p, kf_p = filters.filter_loop(d, prior, R=g["R"], filter_type="b200_parallel")
    assert abs(float(lml_p) - float(g["lml"])) <= TOL * abs(float(g["lml"]))
    assert rel(kf_p['m'], g["mf"]) < TOL and rel(kf_p['P'], g["Pf"]) < TOL
    # host-evaluated transitions (what any prior without on-chip blocks uses) give the same answer
    monkeypatch.setattr(type(pk), "ss_blocks", lambda self: None)
    lml_g, kf_g = filters.filter_loop(d, prior, R=g["R"])
    assert abs(float(lml_g) - float(g["lml"])) <= TOL * abs(float(g["lml"]))
    assert rel(kf_g['P'], g["Pf"]) < TOL
