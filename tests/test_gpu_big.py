"""-m gpu parity tests of the large-block path (libphyss_b200_big.so, BASELINE config 2 shape): separable
spatio-temporal prior Matern-3/2 (time) x RBF (space), state d = 2 Ns, observations m = Ns, against the numpy
oracle; and against the lane-group kernels at a size both cover.  Tolerance 1e-9 relative."""
import numpy as np
import pytest
import torch

from oracle import filters as ofilters
from oracle import sde as osde
from tests import synth
from tests.test_gpu_seq import rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(autouse=True)
def _library_route(monkeypatch):
    """These tests pin the cuBLAS / cuSOLVER-backed library path; separable priors would otherwise take the
    hand-written kernels (tests/test_gpu_kron.py)."""
    from physs_gp_b200 import settings
    monkeypatch.setattr(settings, "kron_kernels", False)


def _st_problem(Ns, T, seed, nan_frac=0.05, irregular=True):
    from physs_gp_b200 import kernels as K
    from physs_gp_b200 import sdes
    rng = np.random.default_rng(seed)
    Xs = rng.uniform(size=(Ns, 2))
    D2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1)
    Ks = np.exp(-0.5 * D2 / 0.2 ** 2) + 1e-6 * np.eye(Ns)             # RBF Gram, lengthscale 0.2
    pprior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(K.Matern32(1.0, 1.0), Ks)]))
    oprior = osde.LTI_SDE([osde.SpaceTimeSeparable(osde.Matern32(1.0, 1.0), Ks)])
    t = synth.time_grid(T, 0.1, rng, irregular=irregular)
    Y = synth.noisy_series(1, T, Ns, rng, nan_frac)[0]
    R = np.tile(0.1 * np.eye(Ns), [T, 1, 1])
    return pprior, oprior, t, Y, R


@pytest.mark.parametrize("Ns,T,irregular", [(20, 40, True), (40, 25, False)])
@pytest.mark.parametrize("jitter", [1e-5, 0.0])
def test_big_block_filter_smoother_match_oracle(cuda_device, Ns, T, irregular, jitter, monkeypatch):
    from physs_gp_b200 import data, filters, settings
    monkeypatch.setattr(settings, "jitter", jitter)
    pprior, oprior, t, Y, R = _st_problem(Ns, T, 3 + Ns, irregular=irregular)
    assert 2 * Ns > settings.big_block_min_dim
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R, jitter)
    d = data.TemporalData(t, Y[:, :, None])
    lml, kf = filters.filter_loop(d, pprior, R=R)
    assert abs(float(lml) - lml_o) <= TOL * abs(lml_o)
    assert rel(kf['m'], mf_o) < TOL and rel(kf['P'], Pf_o) < TOL
    for fs in (False, True):
        ms_o, Ps_o = ofilters.smoother_sequential(oprior, t, mf_o, Pf_o, full_state=fs, jitter=jitter)
        mu, var = filters.smoother_loop(d, pprior, kf, full_state=fs)
        assert rel(mu, ms_o) < TOL and rel(var, Ps_o) < TOL


def test_big_block_equals_lane_group_kernels(cuda_device, monkeypatch):
    """d = 24 (Ns = 12) is covered by both paths: forcing the large-block path must reproduce the
    shared-memory kernels' result."""
    from physs_gp_b200 import data, filters, settings
    pprior, oprior, t, Y, R = _st_problem(12, 200, 9)
    d = data.TemporalData(t, Y[:, :, None])
    lml_a, kf_a = filters.filter_loop(d, pprior, R=R)
    mu_a, var_a = filters.smoother_loop(d, pprior, kf_a, full_state=True)
    monkeypatch.setattr(settings, "big_block_min_dim", 8)
    lml_b, kf_b = filters.filter_loop(d, pprior, R=R)
    mu_b, var_b = filters.smoother_loop(d, pprior, kf_b, full_state=True)
    assert rel(lml_b, lml_a.cpu().numpy()) < TOL
    assert rel(kf_b['P'], kf_a['P'].cpu().numpy()) < TOL and rel(kf_b['m'], kf_a['m'].cpu().numpy()) < TOL
    assert rel(var_b, var_a.cpu().numpy()) < TOL and rel(mu_b, mu_a.cpu().numpy()) < TOL


def test_big_block_non_pd_gives_nan(cuda_device, monkeypatch):
    """a non-PD innovation covariance must come back as NaN (the reference's jnp.linalg.cholesky
    yields NaN, the trainer's NaN guard relies on it), not as a finite partially-factored result."""
    from physs_gp_b200 import data, filters, settings
    pprior, oprior, t, Y, R = _st_problem(20, 12, 5, nan_frac=0.0)
    R = R.copy()
    R[4] = -50.0 * np.eye(20)                                    # S_4 = H P H^T + R_4 is negative definite
    d = data.TemporalData(t, Y[:, :, None])
    lml, kf = filters.filter_loop(d, pprior, R=R)
    assert not np.isfinite(float(lml))
    assert not torch.isfinite(kf['P'][4:]).any() and not torch.isfinite(kf['m'][4:]).any()
    assert torch.isfinite(kf['P'][:4]).all()
