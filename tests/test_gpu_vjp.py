"""-m gpu: reverse pass of the filter's lml (`physs_kf_filter_vjp_f64`) against
oracle/adjoint.py (itself pinned by torch autograd, tests/test_oracle_adjoint.py).  1e-9 relative."""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import adjoint
from oracle import sde as osde

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _close(a, b, what, tol=TOL):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    err = np.abs(a - b).max()
    assert err <= tol * max(np.abs(b).max(), 1e-300), (what, err, np.abs(b).max())


def _given_problem(seed, B, T, d):
    rng = np.random.default_rng(seed)
    A = 0.8 * np.eye(d)[None, None] + 0.1 * rng.normal(size=(B, T, d, d))
    L = rng.normal(size=(B, T, d, d)) * 0.3
    Q = L @ np.swapaxes(L, -1, -2) + 0.05 * np.eye(d)
    H = rng.normal(size=(B, 1, d))
    R = rng.uniform(0.1, 0.5, size=(B, T, 1, 1))
    Y = rng.normal(size=(B, T, 1))
    Y[rng.uniform(size=(B, T, 1)) < 0.15] = np.nan
    m0 = rng.normal(size=(B, d))
    L0 = rng.normal(size=(B, d, d))
    P0 = L0 @ np.swapaxes(L0, -1, -2) + 0.5 * np.eye(d)
    return A, Q, H, R, Y, m0, P0


@pytest.mark.parametrize("time_major", [False, True])
@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_vjp_given_matches_oracle(cuda_device, d, time_major):
    from physs_gp_b200 import ops
    B, T = 5, 37
    A, Q, H, R, Y, m0, P0 = _given_problem(d, B, T, d)
    dev = cuda_device
    tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev)          # noqa: E731
    Yt = tt(Y)
    if time_major:
        Yt = Yt.transpose(0, 1).contiguous().transpose(0, 1)
    dt = torch.zeros((B, T), dtype=torch.float64, device=dev)
    disc = ops.Disc.given(tt(A), tt(Q))
    gl = np.linspace(0.5, 1.5, B)
    lml, mf, Pf = ops.kf_filter(dt, Yt, tt(R), tt(H), tt(m0), tt(P0), disc, jitter=1e-5)
    g = ops.kf_filter_vjp(dt, Yt, tt(R), tt(H), tt(m0), tt(P0), disc, mf, Pf, g_lml=tt(gl), jitter=1e-5,
                          want_R_step=True)
    for b in range(B):
        o = adjoint.filter_lml_vjp(A[b], Q[b], H[b], R[b], Y[b], m0[b][:, None], P0[b], jitter=1e-5, gbar=gl[b])
        assert abs(float(lml[b]) - o["lml"]) <= TOL * abs(o["lml"])
        _close(g["gA"][b], o["gA"], "gA")
        _close(g["gQ"][b], o["gQ"], "gQ")
        _close(g["gH"][b], o["gH"], "gH")
        _close(g["gR_step"][b], o["gR"], "gR_step")
        _close(g["gR"][b], o["gR"].sum(0), "gR")
        _close(g["gm0"][b], o["gm0"][:, 0], "gm0")
        _close(g["gP0"][b], o["gP0"], "gP0")


def _F(lam, s):
    F = np.diag(np.ones(s - 1), 1) if s > 1 else np.zeros((1, 1))
    F[-1] = {1: [-lam], 2: [-lam ** 2, -2 * lam], 3: [-lam ** 3, -3 * lam ** 2, -3 * lam],
             4: [-lam ** 4, -4 * lam ** 3, -6 * lam ** 2, -4 * lam]}[s]
    return F


def _dF(lam, s):
    D = np.zeros((s, s))
    D[-1] = {1: [-1.0], 2: [-2 * lam, -2.0], 3: [-3 * lam ** 2, -6 * lam, -3.0],
             4: [-4 * lam ** 3, -12 * lam ** 2, -12 * lam, -4.0]}[s]
    return D


@pytest.mark.parametrize("d,s", [(1, 1), (2, 2), (2, 1), (3, 3), (4, 4), (4, 2)])
def test_vjp_matern_matches_oracle(cuda_device, d, s):
    from physs_gp_b200 import ops, sdes
    rng = np.random.default_rng(10 * d + s)
    B, T, nblk = 4, 45, d // s
    ls = rng.uniform(0.5, 1.5, size=(B, nblk))
    var = rng.uniform(0.5, 1.5, size=(B, nblk))
    prior = sdes.BatchedMaternSDE(s, ls, var)
    lam, Pinf, H = prior.lam(), prior.P_inf(), prior.H()
    dtn = np.hstack([0.0, rng.uniform(0.05, 0.3, T - 1)])
    Y = rng.normal(size=(B, T, 1))
    Y[rng.uniform(size=(B, T, 1)) < 0.1] = np.nan
    noise = 0.3
    dev = cuda_device
    tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev)          # noqa: E731
    disc = ops.Disc.matern(nblk, tt(lam), tt(Pinf))
    R = torch.full((1, 1, 1, 1), noise, dtype=torch.float64, device=dev)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    lml, mf, Pf = ops.kf_filter(tt(dtn), tt(Y), R, tt(H), m0, tt(Pinf), disc, jitter=1e-5)
    g = ops.kf_filter_vjp(tt(dtn), tt(Y), R, tt(H), m0, tt(Pinf), disc, mf, Pf, jitter=1e-5)
    for b in range(B):
        blocks = lambda x: sla.block_diag(*[sla.expm(_F(l, s) * x) for l in lam[b]])             # noqa: E731
        dA = lambda q, x: (sla.expm_frechet(_F(lam[b, q], s) * x, _dF(lam[b, q], s) * x)[1]       # noqa: E731
                           if x > 0 else np.zeros((s, s)))
        A = np.array([blocks(x) for x in dtn])
        Q = np.array([Pinf[b] - a @ Pinf[b] @ a.T for a in A])
        Rb = np.tile(np.array([[noise]]), [T, 1, 1])
        o = adjoint.filter_lml_vjp(A, Q, H, Rb, Y[b], np.zeros((d, 1)), Pinf[b], jitter=1e-5)
        glam, gPinf = adjoint.matern_chain(blocks, lam[b], dtn, Pinf[b], o["gA"], o["gQ"], dA)
        assert abs(float(lml[b]) - o["lml"]) <= TOL * abs(o["lml"])
        _close(g["glam"][b], glam, "glam", 1e-8)
        _close(g["gPinf"][b], gPinf, "gPinf", 1e-8)
        _close(g["gH"][b], o["gH"], "gH")
        _close(g["gR"][b], o["gR"].sum(0), "gR")
        _close(g["gP0"][b], o["gP0"], "gP0")


def test_model_gradient_matches_finite_differences(cuda_device):
    """SDE_GP.log_marginal_likelihood_and_grad against central differences of the GPU lml itself."""
    from physs_gp_b200 import data, likelihood, models, sdes
    rng = np.random.default_rng(3)
    B, T = 3, 400
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    Y = (np.sin(t)[None] + 0.3 * rng.normal(size=(B, T)))[:, :, None, None]
    ls = rng.uniform(0.6, 1.4, size=(B, 1))
    var = rng.uniform(0.6, 1.4, size=(B, 1))
    noise = 0.2

    def lml_of(ls_, var_, noise_):
        m = models.SDE_GP(data.TemporalData(t, Y), sdes.BatchedMaternSDE(4, ls_, var_), likelihood.Gaussian(noise_))
        return m

    lml, g = lml_of(ls, var, noise).log_marginal_likelihood_and_grad()
    h = 1e-6
    fd_ls = (lml_of(ls + h, var, noise).log_marginal_likelihood()
             - lml_of(ls - h, var, noise).log_marginal_likelihood()) / (2 * h)
    fd_var = (lml_of(ls, var + h, noise).log_marginal_likelihood()
              - lml_of(ls, var - h, noise).log_marginal_likelihood()) / (2 * h)
    fd_n = (lml_of(ls, var, noise + h).log_marginal_likelihood()
            - lml_of(ls, var, noise - h).log_marginal_likelihood()) / (2 * h)
    for got, fd in ((g['lengthscale'][:, 0], fd_ls), (g['variance'][:, 0], fd_var), (g['noise'], fd_n)):
        assert float((got - fd).abs().max()) <= 2e-5 * max(float(fd.abs().max()), 1.0), (got, fd)


def test_vjp_rejects_unsupported_shapes(cuda_device):
    from physs_gp_b200 import _lib
    lib = _lib.load()
    assert lib.physs_kf_vjp_supported(4, 1, _lib.DISC_MATERN, 1) == 1
    assert lib.physs_kf_vjp_supported(8, 1, _lib.DISC_MATERN, 2) == 0
    assert lib.physs_kf_vjp_supported(4, 2, _lib.DISC_GIVEN, 0) == 1      # general (d, m): lane-group kernel
    assert lib.physs_kf_vjp_supported(12, 12, _lib.DISC_GIVEN, 0) == 1
    assert lib.physs_kf_vjp_supported(40, 1, _lib.DISC_GIVEN, 0) == 0
    assert lib.physs_kf_vjp_supported(4, 2, _lib.DISC_MATERN, 1) == 0


def _given_problem_mm(seed, B, T, d, m):
    rng = np.random.default_rng(seed)
    A = 0.8 * np.eye(d)[None, None] + 0.1 * rng.normal(size=(B, T, d, d))
    L = rng.normal(size=(B, T, d, d)) * 0.3
    Q = L @ np.swapaxes(L, -1, -2) + 0.05 * np.eye(d)
    H = np.tile(np.eye(d)[None], [B, 1, 1]) if m == d else rng.normal(size=(B, m, d))
    Lr = rng.normal(size=(B, T, m, m)) * 0.2
    R = Lr @ np.swapaxes(Lr, -1, -2) + 0.2 * np.eye(m)
    Y = rng.normal(size=(B, T, m))
    Y[rng.uniform(size=(B, T, m)) < 0.15] = np.nan                      # partially missing steps
    Y[:, 3] = np.nan                                                    # a fully missing step
    m0 = rng.normal(size=(B, d))
    L0 = rng.normal(size=(B, d, d))
    P0 = L0 @ np.swapaxes(L0, -1, -2) + 0.5 * np.eye(d)
    return A, Q, H, R, Y, m0, P0


@pytest.mark.parametrize("d,m,time_major", [(6, 6, False), (8, 8, True), (12, 12, False), (8, 3, False), (5, 2, True),
                                            (4, 4, False), (3, 2, False), (20, 5, False)])
def test_vjp_general_dims_match_oracle(cuda_device, d, m, time_major):
    """The lane-group reverse pass (csrc/physs_vjp_grp.cu): full-state sites m = d at d = 6 .. 12 -- the shapes the
    reference's VB_NG_ADAM epochs differentiate through -- general H, dense R_k, partially and fully missing steps."""
    from physs_gp_b200 import ops
    B, T = 5, 23
    A, Q, H, R, Y, m0, P0 = _given_problem_mm(100 + d + m, B, T, d, m)
    dev = cuda_device
    tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev)          # noqa: E731
    Yt = tt(Y)
    if time_major:
        Yt = Yt.transpose(0, 1).contiguous().transpose(0, 1)
    dt = torch.zeros((B, T), dtype=torch.float64, device=dev)
    disc = ops.Disc.given(tt(A), tt(Q))
    gl = np.linspace(0.5, 1.5, B)
    Ht = None if m == d else tt(H)
    lml, mf, Pf = ops.kf_filter(dt, Yt, tt(R), Ht, tt(m0), tt(P0), disc, jitter=1e-5)
    g = ops.kf_filter_vjp(dt, Yt, tt(R), Ht, tt(m0), tt(P0), disc, mf, Pf, g_lml=tt(gl), jitter=1e-5,
                          want_R_step=True)
    for b in range(B):
        o = adjoint.filter_lml_vjp(A[b], Q[b], H[b], R[b], Y[b], m0[b][:, None], P0[b], jitter=1e-5, gbar=gl[b])
        assert abs(float(lml[b]) - o["lml"]) <= TOL * abs(o["lml"])
        _close(g["gA"][b], o["gA"], "gA")
        _close(g["gQ"][b], o["gQ"], "gQ")
        _close(g["gH"][b], o["gH"], "gH")
        _close(g["gR_step"][b], o["gR"], "gR_step")
        _close(g["gR"][b], o["gR"].sum(0), "gR")
        _close(g["gm0"][b], o["gm0"][:, 0], "gm0")
        _close(g["gP0"][b], o["gP0"], "gP0")


@pytest.mark.parametrize("case", ["sum_m52_d6", "full_state_m52_d6", "m72x3_d12_independent"])
def test_model_gradient_general_dims_matches_finite_differences(cuda_device, case):
    """SDE_GP.log_marginal_likelihood_and_grad beyond d <= 4, m = 1 (lane-group reverse pass + torch chain through
    expm(F dt) and the closed-form Pinf) against central differences of the GPU lml."""
    from physs_gp_b200 import data, likelihood, models, sdes
    rng = np.random.default_rng(5)
    B, T = 2, 120
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    if case == "sum_m52_d6":
        s, nblk, kw, m = 3, 2, dict(sum_blocks=True), 1
    elif case == "full_state_m52_d6":
        s, nblk, kw, m = 3, 2, dict(full_state_obs=True), 6
    else:
        s, nblk, kw, m = 4, 3, dict(sum_blocks=False), 3
    Y = (np.sin(t)[None, :, None] * np.linspace(0.5, 1.5, m)[None, None] + 0.3 * rng.normal(size=(B, T, m)))[..., None]
    Y[rng.uniform(size=Y.shape) < 0.1] = np.nan
    ls = rng.uniform(0.6, 1.4, size=(B, nblk))
    var = rng.uniform(0.6, 1.4, size=(B, nblk))
    if case == "full_state_m52_d6":
        V = synth_spd(rng, (B, T), m)
        lik = lambda: likelihood.BlockDiagonalGaussian(torch.as_tensor(V, device=cuda_device))     # noqa: E731
    else:
        lik = lambda: likelihood.Gaussian(0.2)                                                      # noqa: E731

    def model(ls_, var_):
        return models.SDE_GP(data.TemporalData(t, Y), sdes.BatchedMaternSDE(s, ls_, var_, **kw), lik())
    lml, g = model(ls, var).log_marginal_likelihood_and_grad()
    h = 1e-6
    for name, arr in (("lengthscale", ls), ("variance", var)):
        for q in range(nblk):
            e = np.zeros_like(arr); e[:, q] = h
            args_p = (ls + e, var) if name == "lengthscale" else (ls, var + e)
            args_m = (ls - e, var) if name == "lengthscale" else (ls, var - e)
            fd = (model(*args_p).log_marginal_likelihood() - model(*args_m).log_marginal_likelihood()) / (2 * h)
            got = g[name][:, q]
            assert float((got - fd).abs().max()) <= 5e-5 * max(float(fd.abs().max()), 1.0), (case, name, q, got, fd)


def synth_spd(rng, lead, m):
    G = rng.normal(size=tuple(lead) + (m, m)) * 0.2
    return G @ np.swapaxes(G, -1, -2) + 0.3 * np.eye(m)
