"""-m gpu parity tests of the hand-written large-block kernels for separable spatio-temporal priors
(physs_kf_filter_kron_f64 / physs_rts_smooth_kron_f64, csrc/physs_kron.cu; BASELINE config 2 shape): persistent
cooperative filter, time-parallel gain kernel + cooperative smoother recursion, against the numpy oracle (dense
Kronecker matrices, as the reference builds them) and against the cuBLAS / cuSOLVER library path.  1e-9 relative."""
import numpy as np
import pytest
import torch

from oracle import filters as ofilters
from oracle import sde as osde
from tests import synth
from tests.test_gpu_seq import rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _problem(Ns, T, seed, kind="m32", nan_frac=0.05, irregular=True, dense_R=False):
    from physs_gp_b200 import kernels as K
    from physs_gp_b200 import sdes
    rng = np.random.default_rng(seed)
    Xs = rng.uniform(size=(Ns, 2))
    D2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1)
    Ks = np.exp(-0.5 * D2 / 0.2 ** 2) + 1e-6 * np.eye(Ns)
    pk = {"m32": K.Matern32, "m52": K.Matern52, "m72": K.Matern72}[kind](1.0, 1.0)
    ok = {"m32": osde.Matern32, "m52": osde.Matern52, "m72": osde.Matern72}[kind](1.0, 1.0)
    pprior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(pk, Ks)]))
    oprior = osde.LTI_SDE([osde.SpaceTimeSeparable(ok, Ks)])
    t = synth.time_grid(T, 0.1, rng, irregular=irregular)
    Y = synth.noisy_series(1, T, Ns, rng, nan_frac)[0]
    R = synth.random_spd(rng, (T,), Ns) if dense_R else np.tile(0.1 * np.eye(Ns), [T, 1, 1])
    return pprior, oprior, t, Y, R


def _run_and_compare(pprior, oprior, t, Y, R, jitter):
    from physs_gp_b200 import data, filters
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R, jitter)
    d = data.TemporalData(t, Y[:, :, None])
    lml, kf = filters.filter_loop(d, pprior, R=R)
    assert rel(kf['P'][0], Pf_o[0]) < TOL, "first filter step"
    assert rel(kf['m'], mf_o) < TOL and rel(kf['P'], Pf_o) < TOL
    assert abs(float(lml) - lml_o) <= TOL * abs(lml_o)
    for fs in (True, False):
        ms_o, Ps_o = ofilters.smoother_sequential(oprior, t, mf_o, Pf_o, full_state=fs, jitter=jitter)
        mu, var = filters.smoother_loop(d, pprior, kf, full_state=fs)
        assert rel(var[-2:], Ps_o[-2:]) < TOL, "last smoother steps (full_state=%s)" % fs
        assert rel(mu, ms_o) < TOL and rel(var, Ps_o) < TOL


@pytest.fixture(autouse=True)
def _kron_on(monkeypatch):
    from physs_gp_b200 import ops, settings
    monkeypatch.setattr(settings, "kron_kernels", True)
    calls = {"f": 0, "s": 0}
    f0, s0 = ops.kf_filter_kron, ops.rts_smooth_kron
    monkeypatch.setattr(ops, "kf_filter_kron", lambda *a, **k: (calls.__setitem__("f", calls["f"] + 1), f0(*a, **k))[1])
    monkeypatch.setattr(ops, "rts_smooth_kron", lambda *a, **k: (calls.__setitem__("s", calls["s"] + 1), s0(*a, **k))[1])
    yield calls


@pytest.mark.parametrize("Ns,T,irregular", [(20, 40, True), (40, 25, False), (37, 12, True), (70, 10, True)])
@pytest.mark.parametrize("jitter", [1e-5, 0.0])
def test_kron_filter_smoother_match_oracle(cuda_device, _kron_on, Ns, T, irregular, jitter, monkeypatch):
    from physs_gp_b200 import settings
    monkeypatch.setattr(settings, "jitter", jitter)
    _run_and_compare(*_problem(Ns, T, 3 + Ns, irregular=irregular), jitter)
    assert _kron_on["f"] == 1 and _kron_on["s"] == 2            # the hand-written route really ran


@pytest.mark.parametrize("kind,Ns", [("m52", 15), ("m72", 12), ("m52", 24)])
def test_kron_other_temporal_kernels(cuda_device, _kron_on, kind, Ns, monkeypatch):
    """ds = 3 / 4 temporal blocks (d = 45 is odd: unaligned rows exercise the non-vector tile loads)."""
    from physs_gp_b200 import settings
    monkeypatch.setattr(settings, "jitter", 1e-5)
    _run_and_compare(*_problem(Ns, 15, 11, kind=kind), 1e-5)
    assert _kron_on["f"] == 1


def test_kron_dense_site_covariance(cuda_device, monkeypatch):
    """Full time-varying R_k (the CVI site covariance of config 2) and no missing data."""
    from physs_gp_b200 import settings
    monkeypatch.setattr(settings, "jitter", 1e-5)
    _run_and_compare(*_problem(48, 8, 5, nan_frac=0.0, dense_R=True), 1e-5)


def test_kron_smoother_crosses_time_chunks(cuda_device, monkeypatch):
    """T larger than one smoother chunk (4 x the CTA count of the gain kernel): the recursion state is handed
    from one cooperative launch to the next."""
    from physs_gp_b200 import settings
    monkeypatch.setattr(settings, "jitter", 1e-5)
    _run_and_compare(*_problem(17, 1300, 21), 1e-5)


def test_kron_equals_library_path(cuda_device, monkeypatch):
    from physs_gp_b200 import data, filters, settings
    pprior, oprior, t, Y, R = _problem(40, 30, 9, dense_R=True)
    d = data.TemporalData(t, Y[:, :, None])
    lml_a, kf_a = filters.filter_loop(d, pprior, R=R)
    mu_a, var_a = filters.smoother_loop(d, pprior, kf_a, full_state=False)
    monkeypatch.setattr(settings, "kron_kernels", False)
    lml_b, kf_b = filters.filter_loop(d, pprior, R=R)
    mu_b, var_b = filters.smoother_loop(d, pprior, kf_b, full_state=False)
    assert rel(lml_a, lml_b.cpu().numpy()) < TOL
    assert rel(kf_a['P'], kf_b['P'].cpu().numpy()) < TOL and rel(kf_a['m'], kf_b['m'].cpu().numpy()) < TOL
    assert rel(var_a, var_b.cpu().numpy()) < TOL and rel(mu_a, mu_b.cpu().numpy()) < TOL


def test_kron_non_pd_gives_nan(cuda_device, monkeypatch):
    from physs_gp_b200 import data, filters
    pprior, oprior, t, Y, R = _problem(20, 12, 5, nan_frac=0.0)
    R = R.copy()
    R[4] = -50.0 * np.eye(20)
    d = data.TemporalData(t, Y[:, :, None])
    lml, kf = filters.filter_loop(d, pprior, R=R)
    assert not np.isfinite(float(lml))
    assert not torch.isfinite(kf['P'][4:]).any() and not torch.isfinite(kf['m'][4:]).any()
    assert torch.isfinite(kf['P'][:4]).all()


# ----------------------------------------------------------------- CVI on large site blocks (config-2 CVI step)
@pytest.mark.parametrize("D,diag", [(40, False), (40, True), (75, False), (200, True)])
def test_big_block_site_update_matches_oracle(cuda_device, D, diag):
    """physs_cvi_natgrad_big_f64 (theta -> lambda, cvi_block_update, lambda -> theta on D x D blocks, one CTA per block)
    against oracle/cvi.py:cvi_step -- itself pinned to the reference's cvi_block_update / theta <-> lambda."""
    from oracle import cvi as ocvi
    from physs_gp_b200 import cvi
    rng = np.random.default_rng(D + diag)
    T = 5 if D < 100 else 3
    Yt = rng.normal(size=(T, D))
    Vt = synth.random_spd(rng, (T,), D, base=0.5, spread=0.1)
    qm = rng.normal(size=(T, D))
    qS = synth.random_spd(rng, (T,), D, base=0.2, spread=0.05)
    dm = rng.normal(size=(T, D))
    if diag:
        dSd = -rng.uniform(0.5, 3.0, size=(T, D))
        dS = np.stack([np.diag(x) for x in dSd])
    else:
        dS = -synth.random_spd(rng, (T,), D, base=0.5, spread=0.1)
    Yo, Vo = ocvi.cvi_step(Yt, Vt, qm, qS, dm, dS, 0.3, ng_jitter=1e-7)
    tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=cuda_device)      # noqa: E731
    Yn, Vn = cvi.natgrad_step(tt(Yt), tt(Vt), tt(qm), tt(qS), None, None, None, 0.3, ng_jitter=1e-7,
                              dm=tt(dm), dS=tt(dSd if diag else dS))
    assert rel(Yn, Yo) < TOL and rel(Vn, Vo) < TOL


def test_big_block_surrogate_ell_matches_oracle(cuda_device):
    from oracle import cvi as ocvi
    from physs_gp_b200 import cvi
    rng = np.random.default_rng(3)
    T, D = 4, 56
    Yt = rng.normal(size=(T, D)); qm = rng.normal(size=(T, D))
    Vt = synth.random_spd(rng, (T,), D, base=0.5, spread=0.1)
    qS = synth.random_spd(rng, (T,), D, base=0.2, spread=0.05)
    tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=cuda_device)      # noqa: E731
    ell = cvi.expected_log_likelihood(tt(qm), tt(qS), tt(Yt), None, cvi.GaussianLik(np.eye(D)), noise=tt(Vt))
    ref = np.array([ocvi.full_gaussian_ell(Yt[t][:, None], Vt[t], qm[t][:, None], qS[t]) for t in range(T)])
    assert rel(ell, ref) < TOL


def test_config2_cvi_iterations_match_oracle(cuda_device, monkeypatch):
    """The config-2 CVI step end to end at a small size: separable Matern-3/2 x RBF prior over Ns = 36 points, ONE
    D = 36 site block per time step, Gaussian likelihood with missing data -- two natural-gradient iterations
    (kron filter + smoother, site update on the large-block kernels) and the ELBO against the numpy oracle."""
    from oracle import cvi as ocvi
    from physs_gp_b200 import cvi, settings
    monkeypatch.setattr(settings, "jitter", 1e-5)
    Ns, T, beta, s2 = 36, 14, 0.5, 0.3
    pprior, oprior, t, Y, _ = _problem(Ns, T, 31)
    q = cvi.FullConjugateGaussian(t, pprior, Ns, B=1)
    model = cvi.VGP(Y[None], cvi.GaussianLik(s2 * np.eye(Ns)), q)
    for _ in range(2):
        model.natural_gradient_update(beta)
    elbo = model.elbo()
    torch.cuda.synchronize()
    Yt = 1e-5 * np.ones((T, Ns)); Vt = np.tile(np.eye(Ns), [T, 1, 1])
    R = s2 * np.eye(Ns)
    for _ in range(2):
        _, qm, qv = ofilters.filter_and_smooth(oprior, t, Yt, Vt)
        dm = np.zeros((T, Ns)); dS = np.zeros((T, Ns, Ns))
        for i in range(T):
            _, dm[i], dS[i] = ocvi.gaussian_ell_and_grads(Y[i], R, np.eye(Ns), qm[i][:, 0], qv[i])
        Yt, Vt = ocvi.cvi_step(Yt, Vt, qm[:, :, 0], qv, dm, dS, beta)
    lml, qm, qv = ofilters.filter_and_smooth(oprior, t, Yt, Vt)
    ell = sum(ocvi.gaussian_ell_and_grads(Y[i], R, np.eye(Ns), qm[i][:, 0], qv[i])[0] for i in range(T))
    ref = ocvi.elbo(ell, ocvi.surrogate_ell(Yt, Vt, qm[:, :, 0], qv), lml)
    assert rel(model.q.Y_tilde[0], Yt) < 1e-8 and rel(model.q.V_tilde[0], Vt) < 1e-8
    assert abs(float(elbo[0]) - ref) <= 1e-8 * abs(ref)


# ----------------------------------------------------------- against REFERENCE output (tests/golden/make_golden_st.py)
import glob  # noqa: E402
import os  # noqa: E402

_ST = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "st_*.npz")))


@pytest.mark.parametrize("path", _ST, ids=lambda p: os.path.basename(p)[:-4])
def test_kron_kernels_match_reference_vectors(cuda_device, _kron_on, path, monkeypatch):
    """The separable-prior kernels against vectors the reference's own SpatioTemporalSeperableKernel + filter_loop +
    smoother_loop produced (no oracle in between): d = 36 (ds = 2, Ns = 18 and ds = 3, Ns = 12), irregular grid,
    dense R_k, partially missing steps, both jitters."""
    from physs_gp_b200 import data, filters, kernels as K, sdes, settings
    g = np.load(path)
    jit = float(g["jitter"])
    monkeypatch.setattr(settings, "jitter", jit)
    kind = K.Matern32 if "m32" in path else K.Matern52
    prior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(kind(*g["temporal"]), g["Ks"])]))
    assert rel(prior.P_inf(), g["P_inf"]) < 1e-14 and np.array_equal(prior.H(), g["H"])
    d = data.TemporalData(g["t"], g["Y"][:, :, None])
    lml, kf = filters.filter_loop(d, prior, R=g["R"])
    assert _kron_on["f"] == 1
    assert abs(float(lml) - float(g["seq_lml"])) <= TOL * abs(float(g["seq_lml"]))
    assert rel(kf['m'], g["seq_mf"]) < TOL and rel(kf['P'], g["seq_Pf"]) < TOL
    for fs in (False, True):
        mu, var = filters.smoother_loop(d, prior, kf, full_state=fs)
        assert rel(mu, g["seq_ms_full%d" % fs]) < TOL and rel(var, g["seq_Ps_full%d" % fs]) < TOL


def test_kron_beyond_resident_size_and_single_step(cuda_device, _kron_on, monkeypatch):
    """m = 212 > 208: the innovation Cholesky and the gain solves no longer fit in shared memory and run the blocked
    out-of-L2 forms (chol_blocked / trsm_fwd / trsm_bwd); and T = 1 (no smoother recursion, no fused predict)."""
    from physs_gp_b200 import settings
    monkeypatch.setattr(settings, "jitter", 1e-5)
    _run_and_compare(*_problem(212, 3, 77), 1e-5)
    _run_and_compare(*_problem(20, 1, 78), 1e-5)
    assert _kron_on["f"] == 2
