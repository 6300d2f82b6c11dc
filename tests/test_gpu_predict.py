"""-m gpu: prediction at new times (`SDE_GP.predict_f`, mirror of stgp/models/sde_gp.py:392-488) -- SURVEY.md
section 8 row f3.  Checked against (a) the dense GP predictive (jitter 0) and (b) the CPU oracle run on the
merged train + test grid (jitter 1e-5)."""
import numpy as np
import pytest
import torch

from oracle import filters as ofilters
from oracle import sde as osde

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _problem(seed=0, T=200, NS=60):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    y = np.sin(t) + 0.3 * rng.normal(size=T)
    y[rng.uniform(size=T) < 0.05] = np.nan
    ts = np.concatenate([rng.uniform(t[0] - 1.0, t[-1] + 1.0, NS - 10), t[rng.integers(0, T, 10)]])
    rng.shuffle(ts)
    return t, y, ts


def _model(t, y, noise, filter_type='b200'):
    from physs_gp_b200 import data, kernels, likelihood, models, sdes
    prior = sdes.LTI_SDE(sdes.Independent([kernels.Matern52(0.8, 0.7)]))
    d = data.TemporalData(t, y[:, None, None] if y.ndim == 1 else y)
    return models.SDE_GP(d, prior, likelihood.Gaussian(noise), filter_type=filter_type)


def test_predict_matches_dense_gp(cuda_device, monkeypatch):
    from physs_gp_b200 import settings
    monkeypatch.setattr(settings, "jitter", 0.0)
    t, y, ts = _problem()
    noise = 0.1
    mu, var = _model(t, y, noise).predict_f(ts)
    assert tuple(mu.shape) == (ts.size, 1, 1) and tuple(var.shape) == (ts.size, 1, 1, 1)
    k = osde.Matern52(0.8, 0.7)
    ob = ~np.isnan(y)
    Kxx = k.K(t[ob], t[ob]) + noise * np.eye(ob.sum())
    Ksx = k.K(ts, t[ob])
    pm = Ksx @ np.linalg.solve(Kxx, y[ob])
    pv = np.diag(k.K(ts, ts)) - np.einsum('ij,ji->i', Ksx, np.linalg.solve(Kxx, Ksx.T))
    np.testing.assert_allclose(mu[:, 0, 0].cpu().numpy(), pm, rtol=0, atol=1e-9)
    np.testing.assert_allclose(var[:, 0, 0, 0].cpu().numpy(), pv, rtol=0, atol=1e-9)


@pytest.mark.parametrize("filter_only", [False, True])
def test_predict_matches_oracle_on_merged_grid(cuda_device, filter_only):
    t, y, ts = _problem(seed=3)
    noise = 0.2
    mdl = _model(t, y, noise)
    mu, var = mdl.predict_f(ts, diagonal=False, filter_only=filter_only, force_full_state=not filter_only)
    # oracle: the same merge by hand
    stacked = np.concatenate([t, ts])
    tt, ui, ri = np.unique(stacked, return_index=True, return_inverse=True)
    yy = np.concatenate([y, np.full(ts.size, np.nan)])[ui]
    prior = osde.LTI_SDE([osde.Matern52(0.8, 0.7)])
    R = np.tile(np.array([[noise]]), [tt.size, 1, 1])
    lml, mf, Pf, _ = ofilters.filter_sequential(prior, tt, yy[:, None], R, jitter=1e-5)
    if filter_only:
        om, oP = mf, Pf
    else:
        om, oP = ofilters.smoother_sequential(prior, tt, mf, Pf, jitter=1e-5, full_state=True)
    sel = ri.reshape(-1)[t.size:]
    om, oP = om[sel], oP[sel]
    assert tuple(mu.shape) == (ts.size, 3, 1) and tuple(var.shape) == (ts.size, 1, 3, 3)
    assert np.abs(mu[..., 0].cpu().numpy() - om.reshape(ts.size, 3)).max() <= TOL * np.abs(om).max()
    assert np.abs(var[:, 0].cpu().numpy() - oP).max() <= TOL * np.abs(oP).max()


def test_predict_batched_and_parallel_route(cuda_device):
    """A batch of series on a shared grid, through the sequential and the parallel-in-time route."""
    t, y, ts = _problem(seed=5, T=600, NS=40)
    rng = np.random.default_rng(9)
    Yb = np.stack([y, y + 0.1 * rng.normal(size=y.size), -y])[:, :, None, None]      # [B, Nt, P, Ns]
    ref = [_model(t, Yb[b, :, 0, 0], 0.1).predict_f(ts) for b in range(3)]
    mu, var = _model(t, Yb, 0.1).predict_f(ts)
    assert tuple(mu.shape) == (3, ts.size, 1, 1)
    for b in range(3):
        assert torch.allclose(mu[b], ref[b][0], rtol=0, atol=1e-12)
        assert torch.allclose(var[b], ref[b][1], rtol=0, atol=1e-12)
    mu_p, var_p = _model(t, Yb, 0.1, filter_type='b200_parallel').predict_f(ts)
    assert float((mu_p - mu).abs().max()) <= 1e-8 * float(mu.abs().max())
    assert float((var_p - var).abs().max()) <= 1e-8 * float(var.abs().max())


def test_predict_argument_check(cuda_device):
    t, y, ts = _problem()
    with pytest.raises(RuntimeWarning):
        _model(t, y, 0.1).predict_f(ts, filter_only=True, force_full_state=True)
