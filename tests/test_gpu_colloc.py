"""-m gpu parity tests of the collocation (EKF) filter step (`physs_kf_filter_colloc_f64`)
through the reference-shaped host API (sdes.PDE prior -> filter_loop / smoother_loop -> C ABI):
  * against tests/golden/ekf_*.npz -- outputs of the reference's own kf_predict_step(PDE) / rts_step_wrapper(PDE)
    (tests/golden/make_golden_ekf.py), no oracle in between;
  * against oracle.filters.filter_pde_sequential on a ragged batch, both memory layouts.
Tolerance 1e-9 relative (array scale)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import filters as ofilters
from oracle import sde as osde
from tests.test_golden import ekf_residuals

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9
FILES = sorted(glob.glob(os.path.join(GOLD, "ekf_*.npz")) + glob.glob(os.path.join(GOLD, "ekfsys_*.npz")))


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _product_pde(g, oracle_res):
    from physs_gp_b200 import kernels as K
    from physs_gp_b200 import sdes
    kind = {"m32": K.Matern32, "m52": K.Matern52, "m72": K.Matern72}
    hyper = np.atleast_2d(g["hyper"])     # one row per latent (ekfsys_*: LotkaVolterra / Lorenz over 2 / 3 latents)
    parent = sdes.LTI_SDE(sdes.Independent([kind[str(k)](float(ls), float(var)) for k, (ls, var) in zip(g["kernel"], hyper)]))
    res = [sdes.PointResidual(r.w, r.terms, r.forcing) for r in oracle_res]
    bnd = g["boundary"] if "boundary" in g.files else None
    return sdes.PDE(parent, res, psuedo_observations=g["y_pseudo"], boundary_conditions=bnd,
                    observe_data=bool(g["observe_data"]))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_cuda_collocation_filter_matches_reference_vectors(cuda_device, path, monkeypatch):
    from physs_gp_b200 import data, filters, settings
    g = np.load(path)
    monkeypatch.setattr(settings, "jitter", float(g["jitter"]))
    prior = _product_pde(g, ekf_residuals(g))
    d = data.TemporalData(g["t"], g["Y"][:, :, None])
    lml, kf = filters.filter_loop(d, prior, R=g["R"])
    if g["kernel"].size > 1:
        # systems: the supplied-transition route (dense A_k, Q_k from the host mirror) runs other instantiations
        monkeypatch.setattr(type(prior), "ss_blocks", lambda self: None)
        lml_g, kf_g = filters.filter_loop(d, prior, R=g["R"])
        assert abs(float(lml_g) - float(g["lml"])) <= TOL * abs(float(g["lml"]))
        assert rel(kf_g['m'], g["mf"]) < TOL and rel(kf_g['P'], g["Pf"]) < TOL
    assert abs(float(lml) - float(g["lml"])) <= TOL * abs(float(g["lml"]))
    assert rel(kf['m'], g["mf"]) < TOL and rel(kf['P'], g["Pf"]) < TOL
    mu, var = filters.smoother_loop(d, prior.parent, kf, full_state=True)
    assert rel(mu, g["ms"]) < TOL and rel(var, g["Ps"]) < TOL


@pytest.mark.parametrize("time_major", [False, True])
def test_cuda_collocation_filter_batch_matches_oracle(cuda_device, time_major):
    """Ragged batch (B = 37) of damped-oscillator series with per-series data, both memory layouts, DISC_MATERN."""
    from physs_gp_b200 import ops
    rng = np.random.default_rng(17)
    B, T, d = 37, 60, 4
    kern = osde.Matern72(0.6, 2.0)
    prior = osde.LTI_SDE([kern])
    a_, b_ = 9.81 / 1.3, 0.35
    res = [ofilters.PointResidual([0.0, b_, 1.0, 0.0], [("sin", 0, a_)]),
           ofilters.PointResidual([0.0, a_, b_, 1.0], [], forcing=0.05 * np.sin(np.arange(T)))]
    t = np.cumsum(rng.uniform(0.5, 1.5, T) * 0.05)
    Y = 0.8 * np.cos(2.5 * t)[None, :, None] + 0.05 * rng.normal(size=(B, T, 1))
    Y[rng.uniform(size=Y.shape) < 0.3] = np.nan
    R = np.tile(0.05 ** 2 * np.eye(1), [T, 1, 1])
    bnd = np.full((B, T, 1), np.nan)
    bnd[:, 0, 0] = 0.8 + 0.01 * rng.normal(size=B)
    dev = cuda_device
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    Yd, bd = tt(Y), tt(bnd)
    if time_major:
        Yd = Yd.transpose(0, 1).contiguous().transpose(0, 1)
    lam = tt(np.array([[np.sqrt(7.0) / 0.6]]))
    Pinf = tt(prior.P_inf()[None])
    disc = ops.Disc.matern(1, lam, Pinf)
    dt = tt(np.hstack([0.0, np.diff(t)]))
    terms = [(p, k, i, c) for p, r in enumerate(res) for k, i, c in r.terms]
    forcing = tt(np.stack([np.zeros(T), res[1].forcing]))
    lml, mf, Pf, lk = ops.kf_filter_colloc(dt, Yd, tt(R[None]), tt(prior.H()[None]), tt(np.zeros((1, d))), Pinf, disc,
                                           np.stack([r.w for r in res]), terms, forcing=forcing, y_pseudo=[0.0, 0.0],
                                           boundary=bd, observe_data=True, jitter=1e-5, want_lml_k=True)
    for b in (0, 5, 36):
        lml_o, mf_o, Pf_o, lk_o = ofilters.filter_pde_sequential(prior, res, t, Y[b], R, boundary=bnd[b],
                                                                 y_pseudo=[0.0, 0.0], observe_data=True, jitter=1e-5)
        assert abs(float(lml[b]) - lml_o) <= TOL * abs(lml_o)
        assert rel(mf[b], mf_o[..., 0]) < TOL and rel(Pf[b], Pf_o) < TOL
        assert rel(lk[b], lk_o) < TOL
