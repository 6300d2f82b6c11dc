"""-m gpu parity tests for the shared-memory (lane-group) filter / smoother used for state dims > 4,
through the same host API, against the C oracle (validated against the numpy oracle in
tests/test_c_oracle.py).  Tolerance 1e-9 relative (max-abs over the array scale), see test_gpu_seq."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import filters as ofilters
from oracle import sde as osde
from tests import synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


CASES = [
    # (block size s, nblk, H kind, B, T)
    (3, 2, "indep", 5, 160),      # d=6, m=2   (two Matern-5/2 latents)
    (3, 2, "full", 5, 160),       # d=6, m=6   (config 3: derivative-augmented full-state sites)
    (4, 2, "sum", 9, 200),        # d=8, m=1   (config 5, d=8)
    (4, 2, "full", 3, 120),       # d=8, m=8
    (4, 3, "full", 3, 100),       # d=12, m=12 (config 3)
    (4, 3, "indep", 4, 100),      # d=12, m=3
    (4, 4, "sum", 5, 120),        # d=16, m=1  (config 5)
    (2, 5, "sum", 4, 100),        # d=10, m=1
    (4, 4, "full", 3, 60),        # d=16, m=16 (compile-time shape variant m == d at DM = 16)
    (4, 2, "indep", 4, 90),       # d=8, m=2   (compile-time d and block size, runtime m)
    (4, 8, "sum", 3, 80),         # d=32, m=1  (config 5)
    (4, 8, "full", 2, 40),        # d=32, m=32
    (1, 5, "indep", 3, 50),       # d=5 Ornstein-Uhlenbeck blocks
]


def _H(kind, s, nblk):
    d = s * nblk
    if kind == "full":
        return np.eye(d)
    if kind == "sum":
        h = np.zeros([1, d])
        h[0, ::s] = 1.0
        return h
    H = np.zeros([nblk, d])
    for b in range(nblk):
        H[b, b * s] = 1.0
    return H


@pytest.mark.parametrize("s,nblk,hkind,B,T", CASES)
@pytest.mark.parametrize("full_state", [True, False])
def test_group_kernels_match_oracle(cuda_device, s, nblk, hkind, B, T, full_state):
    from physs_gp_b200 import data, filters, sdes
    rng = np.random.default_rng(1000 * s + 10 * nblk + len(hkind))
    ls = synth.log_uniform(rng, 0.5, 2.0, (B, nblk))
    var = synth.log_uniform(rng, 0.5, 2.0, (B, nblk))
    prior = sdes.BatchedMaternSDE(s, ls, var, sum_blocks=(hkind == "sum"), full_state_obs=(hkind == "full"))
    H = _H(hkind, s, nblk)
    assert np.array_equal(H, prior.H())
    m = H.shape[0]
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(B, T, m, rng, 0.08)
    R = synth.random_spd(rng, (B, T), m)
    d_ = data.TemporalData(t, Y[..., None])
    lml, kf = filters.filter_loop(d_, prior, R=R)
    mu, var_s = filters.smoother_loop(d_, prior, kf, full_state=full_state)
    torch.cuda.synchronize()
    ref = c_oracle.filter_smooth(s, prior.lam(), prior.P_inf(), H, t, Y, R, jitter=1e-5,
                                 full_state=full_state)
    assert rel(lml, ref["lml"]) < TOL
    assert rel(kf["m"][..., 0], ref["mf"]) < TOL
    assert rel(kf["P"], ref["Pf"]) < TOL
    assert rel(mu[..., 0], ref["ms"]) < TOL
    assert rel(var_s, ref["Ps"]) < TOL


def test_group_given_mode_generic_prior(cuda_device):
    """d = 5 dense LTI prior with scipy expm -> DISC_GIVEN through the group kernels."""
    import scipy.linalg as sla
    from physs_gp_b200 import data, filters
    rng = np.random.default_rng(77)
    d = 5
    Mx = rng.normal(size=(d, d))
    F = -(Mx @ Mx.T) * 0.3 - 0.5 * np.eye(d) + 0.4 * (Mx - Mx.T)
    Pinf = sla.solve_continuous_lyapunov(F, -np.eye(d))
    Hm = np.zeros([2, d]); Hm[0, 0] = 1.0; Hm[1, 3] = 1.0

    class GenericPrior:
        def m_inf(self, x, X_s, t): return np.zeros([d, 1])
        def P_inf(self, x, X_s, t): return Pinf
        def H(self, x, X_s, t): return Hm
        def expm(self, X_s, dt): return sla.expm(F * dt)
        def Q(self, dt, A, P, X_spatial=None): return P - A @ P @ A.T

    oprior = osde.LTI_SDE([osde.GenericLTI(F, Hm, Pinf)])
    T = 120
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(1, T, 2, rng, 0.1)[0]
    R = synth.random_spd(rng, (T,), 2)
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R)
    ms_o, Ps_o = ofilters.smoother_sequential(oprior, t, mf_o, Pf_o)
    d_ = data.TemporalData(t, Y[:, :, None])
    lml, kf = filters.filter_loop(d_, GenericPrior(), R=R)
    mu, var = filters.smoother_loop(d_, GenericPrior(), kf)
    assert abs(float(lml) - lml_o) <= TOL * abs(lml_o)
    assert rel(kf["m"], mf_o) < TOL and rel(kf["P"], Pf_o) < TOL
    assert rel(mu, ms_o) < TOL and rel(var, Ps_o) < TOL


def test_group_all_missing_and_single_step(cuda_device):
    from physs_gp_b200 import data, filters, sdes
    rng = np.random.default_rng(5)
    prior = sdes.BatchedMaternSDE(4, synth.log_uniform(rng, 0.5, 2.0, (3, 2)))
    for T, nan_frac in ((1, 0.0), (2, 0.0), (40, 1.0)):
        t = synth.time_grid(T, 0.1, rng)
        Y = synth.noisy_series(3, T, 1, rng, nan_frac)
        R = np.full([1, 1, 1, 1], 0.1)
        d_ = data.TemporalData(t, Y[..., None])
        lml, kf = filters.filter_loop(d_, prior, R=R)
        mu, var = filters.smoother_loop(d_, prior, kf, full_state=True)
        ref = c_oracle.filter_smooth(4, prior.lam(), prior.P_inf(), prior.H(), t, Y, R[0, 0], full_state=True)
        if nan_frac == 1.0:
            assert float(lml.abs().max()) == 0.0
        else:
            assert rel(lml, ref["lml"]) < TOL
        assert rel(kf["P"], ref["Pf"]) < TOL and rel(var, ref["Ps"]) < TOL and rel(mu[..., 0], ref["ms"]) < TOL

