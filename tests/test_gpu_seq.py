"""-m gpu parity tests: the CUDA filter / smoother, called through the reference-shaped host API
(filter_loop / smoother_loop -> evoke('filter'|'smoother', 'b200') -> C ABI), against the numpy
oracle on identical seeded inputs.

Tolerance (north_star): 1e-9 relative in fp64.  "Relative" is max|a - b| / max|b| per array: the
reference itself mixes O(1) and O(lam^6) state components, so element-wise relative error is not
meaningful for near-zero entries."""
import numpy as np
import pytest
import torch

from oracle import filters as ofilters
from oracle import sde as osde
from tests import synth

pytestmark = pytest.mark.gpu

TOL = 1e-9


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _pair(kind, ls, var):
    """(product kernel, oracle kernel) with identical hyper-parameters."""
    from physs_gp_b200 import kernels as K
    prod = {"m12": None, "m32": K.Matern32, "m52": K.Matern52, "m72": K.Matern72}[kind](ls, var)
    orac = {"m32": osde.Matern32, "m52": osde.Matern52, "m72": osde.Matern72}[kind](ls, var)
    return prod, orac


def _priors(spec, full_state_obs=False, keep_dims=None):
    """spec: list of latents, each a list of (kind, ls, var) summed."""
    from physs_gp_b200 import kernels as K
    from physs_gp_b200 import sdes
    plat, olat = [], []
    for latent in spec:
        pk = [_pair(*p)[0] for p in latent]
        ok = [_pair(*p)[1] for p in latent]
        plat.append(K.sum_kernels(pk))
        olat.append(ok[0] if len(ok) == 1 else osde.SumKernel(ok))
    if full_state_obs:
        return (sdes.LTI_SDE_Full_State_Obs(sdes.Independent(plat), keep_dims=keep_dims),
                osde.LTI_SDE_Full_State_Obs(olat, keep_dims=keep_dims))
    return sdes.LTI_SDE(sdes.Independent(plat)), osde.LTI_SDE(olat)


SPECS = {
    "c1_m32": ([[("m32", 1.0, 1.3)]], False),
    "m52": ([[("m52", 0.7, 0.9)]], False),
    "m72": ([[("m72", 1.2, 1.1)]], False),
    "sum_m32x2": ([[("m32", 1.0, 1.3), ("m32", 0.4, 0.5)]], False),
    "indep_m32x2": ([[("m32", 1.0, 1.3)], [("m32", 0.4, 0.5)]], False),
    "m32_fullstate": ([[("m32", 1.0, 1.3)]], True),
    "m52_fullstate": ([[("m52", 0.7, 0.9)]], True),
    "m72_fullstate": ([[("m72", 1.2, 1.1)]], True),
    "indep_m32x2_fullstate": ([[("m32", 1.0, 1.3)], [("m32", 0.4, 0.5)]], True),
}


def _run_case(spec, fso, T, seed, jitter, nan_frac, full_state, time_varying_R=True, keep_dims=None):
    from physs_gp_b200 import data, filters, settings
    rng = np.random.default_rng(seed)
    pprior, oprior = _priors(spec, fso, keep_dims)
    H = oprior.H()
    m = H.shape[0]
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(1, T, m, rng, nan_frac)[0]
    R = synth.random_spd(rng, (T,), m) if time_varying_R else np.tile(0.1 * np.eye(m), [T, 1, 1])
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R, jitter)
    ms_o, Ps_o = ofilters.smoother_sequential(oprior, t, mf_o, Pf_o, full_state=full_state, jitter=jitter)
    old = settings.jitter
    settings.jitter = jitter
    try:
        d = data.TemporalData(t, Y[:, :, None])
        lml, kf = filters.filter_loop(d, pprior, R=R if time_varying_R else R[:1])
        mu, var = filters.smoother_loop(d, pprior, kf, full_state=full_state)
    finally:
        settings.jitter = old
    torch.cuda.synchronize()
    assert abs(float(lml) - lml_o) <= TOL * abs(lml_o), (float(lml), lml_o)
    assert rel(kf['m'], mf_o) < TOL
    assert rel(kf['P'], Pf_o) < TOL
    assert rel(mu, ms_o) < TOL
    assert rel(var, Ps_o) < TOL


@pytest.mark.parametrize("name", sorted(SPECS))
@pytest.mark.parametrize("jitter", [1e-5, 0.0])
def test_filter_smoother_parity(cuda_device, name, jitter):
    spec, fso = SPECS[name]
    _run_case(spec, fso, T=300, seed=hash(name) % 1000, jitter=jitter, nan_frac=0.08, full_state=False)


@pytest.mark.parametrize("name", ["m72", "sum_m32x2", "m52_fullstate"])
def test_full_state_output(cuda_device, name):
    spec, fso = SPECS[name]
    _run_case(spec, fso, T=200, seed=7, jitter=1e-5, nan_frac=0.05, full_state=True)


def test_keep_dims_partial_state_observation(cuda_device):
    # LTI_SDE_Full_State_Obs_With_Mask (sdes.py:174-190): observe [f, df/dt] of a Matern-7/2 state
    _run_case([[("m72", 1.2, 1.1)]], True, T=200, seed=11, jitter=1e-5, nan_frac=0.05,
              full_state=False, keep_dims=[0, 1])


def test_config1_matern32_T10k(cuda_device):
    """BASELINE config 1: 1D Matern-3/2, Gaussian likelihood, N = 10k, exact lml."""
    _run_case(SPECS["c1_m32"][0], False, T=10000, seed=0, jitter=1e-5, nan_frac=0.05,
              full_state=False, time_varying_R=False)


def test_edge_cases(cuda_device):
    spec, fso = SPECS["m52"]
    _run_case(spec, fso, T=1, seed=3, jitter=1e-5, nan_frac=0.0, full_state=False)      # single step
    _run_case(spec, fso, T=2, seed=4, jitter=1e-5, nan_frac=0.0, full_state=True)
    _run_case(spec, fso, T=50, seed=5, jitter=1e-5, nan_frac=1.0, full_state=False)     # all missing
    _run_case(SPECS["indep_m32x2_fullstate"][0], True, T=60, seed=6, jitter=1e-5, nan_frac=0.5,
              full_state=False)                                                        # ragged masks


def test_generic_prior_goes_through_given_A_Q(cuda_device):
    """A prior with no closed-form block description (dense F, scipy expm) takes the DISC_GIVEN route:
    the shim evaluates prior.expm / prior.Q per distinct dt on the host, as the reference API allows."""
    from physs_gp_b200 import data, filters
    rng = np.random.default_rng(12)
    F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-2.0, -3.0, -1.5]])
    import scipy.linalg as sla
    Pinf = sla.solve_continuous_lyapunov(F, -np.diag([0.0, 0.0, 1.0]))
    Hm = np.array([[1.0, 0.0, 0.0]])
    ok = osde.GenericLTI(F, Hm, Pinf)
    oprior = osde.LTI_SDE([ok])

    class GenericPrior:                      # duck-typed reference prior API
        def m_inf(self, x, X_s, t): return np.zeros([3, 1])
        def P_inf(self, x, X_s, t): return Pinf
        def H(self, x, X_s, t): return Hm
        def expm(self, X_s, dt): return sla.expm(F * dt)
        def Q(self, dt, A, P, X_spatial=None): return P - A @ P @ A.T

    T = 150
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(1, T, 1, rng, 0.05)[0]
    R = np.tile(0.2 * np.eye(1), [T, 1, 1])
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R)
    ms_o, Ps_o = ofilters.smoother_sequential(oprior, t, mf_o, Pf_o)
    d = data.TemporalData(t, Y[:, :, None])
    lml, kf = filters.filter_loop(d, GenericPrior(), R=R)
    mu, var = filters.smoother_loop(d, GenericPrior(), kf)
    assert abs(float(lml) - lml_o) <= TOL * abs(lml_o)
    assert rel(kf['P'], Pf_o) < TOL and rel(mu, ms_o) < TOL and rel(var, Ps_o) < TOL


@pytest.mark.parametrize("time_major", [True, False])
@pytest.mark.parametrize("s,nblk", [(2, 1), (3, 1), (4, 1), (2, 2)])
def test_batched_per_series_hyperparameters(cuda_device, s, nblk, time_major, monkeypatch):
    """B series with their own lengthscales (BASELINE config 5 shape, small): every series must equal
    the oracle run on that series alone -- in both memory orders of the batch."""
    from physs_gp_b200 import data, filters, sdes, settings
    monkeypatch.setattr(settings, "time_major", time_major)
    rng = np.random.default_rng(100 + 10 * s + nblk)
    B, T = 37, 120
    ls = synth.log_uniform(rng, 0.5, 2.0, (B, nblk))
    var = synth.log_uniform(rng, 0.5, 2.0, (B, nblk))
    prior = sdes.BatchedMaternSDE(s, ls, var, sum_blocks=True)
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(B, T, 1, rng, 0.05)
    R = np.full([1, 1, 1, 1], 0.1)
    d = data.TemporalData(t, Y[..., None])
    lml, kf = filters.filter_loop(d, prior, R=R)
    mu, var_s = filters.smoother_loop(d, prior, kf, full_state=True)
    for o in (kf['P'], var_s):
        assert o.transpose(0, 1).is_contiguous() == time_major and o.shape[:2] == (B, T)
    kind = {2: osde.Matern32, 3: osde.Matern52, 4: osde.Matern72}[s]
    for b in range(0, B, 6):
        parts = [kind(ls[b, i], var[b, i]) for i in range(nblk)]
        op = osde.LTI_SDE([parts[0] if nblk == 1 else osde.SumKernel(parts)])
        lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(op, t, Y[b], np.tile(R[0, 0], [T, 1, 1]))
        ms_o, Ps_o = ofilters.smoother_sequential(op, t, mf_o, Pf_o, full_state=True)
        assert abs(float(lml[b]) - lml_o) <= TOL * abs(lml_o)
        assert rel(kf['m'][b], mf_o) < TOL and rel(kf['P'][b], Pf_o) < TOL
        assert rel(mu[b], ms_o) < TOL and rel(var_s[b], Ps_o) < TOL


def test_full_size_properties(cuda_device):
    """Size-independent properties at a BASELINE-sized shard (B = 2048 x T = 10k, Matern-7/2):
    (i) permuting the series permutes the outputs bit-exactly (no cross-series coupling);
    (ii) smoothed variance <= filtered variance; (iii) last smoothed step == last filtered step;
    (iv) an all-missing series reproduces the prior (P = Pinf, lml = 0)."""
    from physs_gp_b200 import data, filters, sdes
    rng = np.random.default_rng(5)
    B, T = 2048, 10000
    ls = synth.log_uniform(rng, 5.0, 20.0, (B, 1)) * 0.1
    prior = sdes.BatchedMaternSDE(4, ls)
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(B, T, 1, rng, 0.05)
    Y[3] = np.nan
    R = np.full([1, 1, 1, 1], 0.1)
    d = data.TemporalData(t, Y[..., None])
    lml, kf = filters.filter_loop(d, prior, R=R)
    mu, var = filters.smoother_loop(d, prior, kf, full_state=True)
    assert torch.isfinite(lml).all() and torch.isfinite(var).all()
    assert float(lml[3]) == 0.0
    Pinf3 = torch.as_tensor(prior.P_inf()[3], device=var.device)
    assert rel(kf['P'][3, -1], Pinf3.cpu().numpy()) < 1e-12
    assert torch.equal(mu[:, -1], kf['m'][:, -1]) and torch.equal(var[:, -1], kf['P'][:, -1])
    dP = torch.diagonal(kf['P'], dim1=-2, dim2=-1) - torch.diagonal(var, dim1=-2, dim2=-1)
    scale = torch.diagonal(kf['P'], dim1=-2, dim2=-1).abs().amax(dim=(0, 1))
    assert (dP / scale >= -1e-9).all()
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    prior_p = sdes.BatchedMaternSDE(4, ls[perm.numpy()])
    d_p = data.TemporalData(t, Y[perm.numpy()][..., None])
    lml_p, kf_p = filters.filter_loop(d_p, prior_p, R=R)
    mu_p, var_p = filters.smoother_loop(d_p, prior_p, kf_p, full_state=True)
    permd = perm.to(lml.device)
    assert torch.equal(lml_p.nan_to_num(), lml[permd].nan_to_num())
    assert torch.equal(var_p, var[permd]) and torch.equal(mu_p, mu[permd])


@pytest.mark.parametrize("d,m,given,mo", [(1, 1, False, 0), (2, 1, False, 0), (3, 1, False, 1), (3, 3, True, 0),
                                          (4, 1, False, 0), (4, 2, False, 2), (4, 4, True, 0), (8, 1, False, 0),
                                          (8, 3, True, 2)])
def test_time_major_equals_batch_major_bitwise(cuda_device, d, m, given, mo):
    """The memory order of the batch must not change a single bit of any series' result (ragged last
    warp: B = 70), for the register kernels (d <= 4) and the shared-memory kernels (d = 8)."""
    from physs_gp_b200 import ops
    dev = cuda_device
    rng = np.random.default_rng(1000 + 10 * d + m)
    B, T = 70, 64
    t = synth.time_grid(T, 0.1, rng)
    dt_f = torch.as_tensor(np.hstack([0.0, np.diff(t)]), device=dev)
    dt_s = torch.as_tensor(np.hstack([np.diff(t), 0.0]), device=dev)
    Y = torch.as_tensor(synth.noisy_series(B, T, m, rng, 0.1), device=dev)
    R = torch.as_tensor(synth.random_spd(rng, (B, T), m), device=dev)
    H = torch.as_tensor(rng.normal(size=(1, m, d)), device=dev)
    if given:
        Fm = rng.normal(size=(d, d)) * 0.3 - np.eye(d)
        import scipy.linalg as sla
        Pinf = sla.solve_continuous_lyapunov(Fm, -np.eye(d))
        A = np.stack([sla.expm(Fm * x) for x in np.hstack([0.0, np.diff(t)])])
        Q = Pinf - A @ Pinf @ np.swapaxes(A, -1, -2)
        A_s = np.stack([sla.expm(Fm * x) for x in np.hstack([np.diff(t), 0.0])])
        Q_s = Pinf - A_s @ Pinf @ np.swapaxes(A_s, -1, -2)
        disc_f = ops.Disc.given(torch.as_tensor(A[None], device=dev), torch.as_tensor(Q[None], device=dev))
        disc_s = ops.Disc.given(torch.as_tensor(A_s[None], device=dev), torch.as_tensor(Q_s[None], device=dev))
        P0 = torch.as_tensor(Pinf[None], device=dev)
    else:
        from physs_gp_b200 import sdes
        s_blk = d if d <= 4 else 4
        prior = sdes.BatchedMaternSDE(s_blk, synth.log_uniform(rng, 0.5, 2.0, (B, d // s_blk)))
        lam = torch.as_tensor(prior.lam(), device=dev)
        P0 = torch.as_tensor(prior.P_inf(), device=dev)
        disc_f = disc_s = ops.Disc.matern(d // s_blk, lam, P0)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    Hout = torch.as_tensor(rng.normal(size=(mo, d)), device=dev) if mo else None
    res = []
    for tm in (False, True):
        Yl = Y.transpose(0, 1).contiguous().transpose(0, 1) if tm else Y
        lml, mf, Pf, lk = ops.kf_filter(dt_f, Yl, R, H, m0, P0, disc_f, jitter=1e-5, want_lml_k=True)
        ms, Ps = ops.rts_smooth(dt_s, mf, Pf, disc_s, Hout=Hout, jitter=1e-5)
        assert Pf.transpose(0, 1).is_contiguous() == tm and Ps.transpose(0, 1).is_contiguous() == tm
        res.append((lml, mf, Pf, lk, ms, Ps))
    torch.cuda.synchronize()
    for a, b in zip(*res):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b)


def test_broadcast_transitions_are_materialised(cuda_device):
    """Disc.given(A, Q) with time-invariant A, Q ([d, d], broadcast over T) and a dt of shape [1]
    must give the same result as the explicit [T, d, d] arrays (the kernels index the time axis densely)."""
    from physs_gp_b200 import ops
    dev = cuda_device
    rng = np.random.default_rng(5)
    B, T, d = 3, 50, 2
    oprior = osde.LTI_SDE([osde.Matern32(0.8, 1.1)])
    A1 = oprior.expm(0.1)
    Q1 = oprior.Q(0.1, A1, oprior.P_inf())
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    Y = tt(rng.normal(size=(B, T, 1)))
    R = tt(np.full((1, 1, 1, 1), 0.1))
    H, m0, P0 = tt(oprior.H()[None]), tt(np.zeros((1, d))), tt(oprior.P_inf()[None])
    full = ops.Disc.given(tt(np.tile(A1, [T, 1, 1])), tt(np.tile(Q1, [T, 1, 1])))
    bcast = ops.Disc.given(tt(A1), tt(Q1))
    dt_full, dt_b = tt(np.full(T, 0.1)), tt(np.full(1, 0.1))
    lml_a, mf_a, Pf_a = ops.kf_filter(dt_full, Y, R, H, m0, P0, full, jitter=1e-5)
    lml_b, mf_b, Pf_b = ops.kf_filter(dt_b, Y, R, H, m0, P0, bcast, jitter=1e-5)
    assert torch.equal(lml_a, lml_b) and torch.equal(mf_a, mf_b) and torch.equal(Pf_a, Pf_b)
    ms_a, Ps_a = ops.rts_smooth(dt_full, mf_a, Pf_a, full, Hout=None, jitter=1e-5)
    ms_b, Ps_b = ops.rts_smooth(dt_b, mf_b, Pf_b, bcast, Hout=None, jitter=1e-5)
    assert torch.equal(ms_a, ms_b) and torch.equal(Ps_a, Ps_b)


def test_scalar_update_non_pd_innovation_gives_nan_lml(cuda_device):
    """with m = 1 two steps with S < 0 must not cancel in the running determinant product --
    the reference's cholesky(S) gives NaN at each such step."""
    from physs_gp_b200 import ops
    dev = cuda_device
    rng = np.random.default_rng(6)
    B, T, d = 2, 20, 2
    oprior = osde.LTI_SDE([osde.Matern32(0.8, 1.1)])
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    Y = tt(rng.normal(size=(B, T, 1)))
    Rn = np.full((1, T, 1, 1), 0.1)
    Rn[0, 3] = Rn[0, 7] = -25.0                                   # two negative innovation variances
    disc = ops.Disc.matern(1, tt(np.full((1, 1), np.sqrt(3.0) / 0.8)), tt(oprior.P_inf()[None]))
    lml, _, _ = ops.kf_filter(tt(np.full(T, 0.1)), Y, tt(Rn), tt(oprior.H()[None]), tt(np.zeros((1, d))),
                              tt(oprior.P_inf()[None]), disc, jitter=1e-5)
    assert torch.isnan(lml).all()


def test_int64_indexing_beyond_2_31_elements(cuda_device):
    """one array with more than 2^31 elements (Pf: 36,864 x 3,700 x 16 = 2.18e9 doubles, 17.5 GB)
    so that every per-step offset above 2^31 goes through the int64 index paths, time-major and batch-major.
    Checked against the same series run as a small batch (bit-for-bit: a series' arithmetic does not depend on
    the batch it sits in) at the last rows of the big arrays, and against the smoother on top."""
    from physs_gp_b200 import ops
    dev = cuda_device
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of free device memory")
    B, T, d = 36864, 3700, 4
    assert B * T * d * d > 2 ** 31
    rng = np.random.default_rng(3)
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    from physs_gp_b200 import sdes
    ls = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1)))
    prior = sdes.BatchedMaternSDE(4, ls)
    lam, Pinf, H = tt(prior.lam()), tt(prior.P_inf()), tt(prior.H())
    dt = tt(np.hstack([0.0, rng.uniform(0.05, 0.15, T - 1)]))
    dts = torch.cat([dt[1:], torch.zeros(1, dtype=torch.float64, device=dev)])
    g = torch.Generator(device=dev).manual_seed(5)
    R = tt(np.full((1, 1, 1, 1), 0.1))
    m0 = tt(np.zeros((1, d)))
    sel = torch.tensor([0, 1, B // 2, B - 33, B - 2, B - 1], device=dev)
    for time_major in (True, False):
        Y = torch.randn((T, B, 1) if time_major else (B, T, 1), generator=g, device=dev, dtype=torch.float64)
        Y = Y.transpose(0, 1) if time_major else Y
        disc = ops.Disc.matern(1, lam, Pinf)
        lml, mf, Pf = ops.kf_filter(dt, Y, R, H, m0, Pinf, disc, jitter=1e-5)
        ms, Ps = ops.rts_smooth(dts, mf, Pf, disc, Hout=None, jitter=1e-5)
        small = ops.Disc.matern(1, lam[sel], Pinf[sel])
        Ys = Y[sel].contiguous()
        lml_s, mf_s, Pf_s = ops.kf_filter(dt, Ys, R, H, m0, Pinf[sel], small, jitter=1e-5)
        ms_s, Ps_s = ops.rts_smooth(dts, mf_s, Pf_s, small, Hout=None, jitter=1e-5)
        assert torch.equal(lml[sel], lml_s)
        assert torch.equal(mf[sel], mf_s) and torch.equal(Pf[sel], Pf_s)
        assert torch.equal(ms[sel], ms_s) and torch.equal(Ps[sel], Ps_s)
        assert bool(torch.isfinite(Ps[-1, 0]).all()) and bool(torch.isfinite(Ps[-1, -1]).all())
        del Y, mf, Pf, ms, Ps
        torch.cuda.empty_cache()


def test_precision_sites_match_covariance_sites(cuda_device):
    """R_inv (precision-parameterised sites) on the sequential b200 path == the same filter fed
    R = inv(R_inv); the reference's own sequential path raises for R_inv (kalman_filter.py:67)."""
    from physs_gp_b200 import data, filters
    rng = np.random.default_rng(12)
    pprior, oprior = _priors([[("m32", 1.0, 1.3)], [("m32", 0.4, 0.5)]], False)
    T, m = 60, 2
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(1, T, m, rng, 0.1)[0]
    R = synth.random_spd(rng, (T,), m)
    Rinv = np.linalg.inv(R)
    d = data.TemporalData(t, Y[:, :, None])
    lml_a, kf_a = filters.filter_loop(d, pprior, R=R)
    lml_b, kf_b = filters.filter_loop(d, pprior, R_inv=Rinv)
    assert abs(float(lml_a) - float(lml_b)) <= 1e-10 * abs(float(lml_a))
    assert rel(kf_b['m'], kf_a['m'].cpu().numpy()) < 1e-10 and rel(kf_b['P'], kf_a['P'].cpu().numpy()) < 1e-10
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R, 1e-5)
    assert rel(kf_b['m'], mf_o) < TOL and rel(kf_b['P'], Pf_o) < TOL


def test_reciprocal_primitives_are_ulp_accurate(cuda_device):
    """fast_rcp / fast_rsqrt (physs_core.cuh: hardware seed + two Newton steps) through the 1 x 1 case of
    physs_spd_inverse_f64, (a)^-1 = (1 * rsqrt(a)) * rsqrt(a): a few ulp over twelve decades -- a seed worse than
    2^-14 (or a missing Newton step) would show up here long before the 1e-9 parity tolerance notices."""
    from physs_gp_b200 import ops
    rng = np.random.default_rng(0)
    a = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), 200000))
    inv = ops.spd_inverse(torch.as_tensor(a, device=cuda_device).reshape(-1, 1, 1), 0.0).reshape(-1).cpu().numpy()
    err = np.abs(inv * a - 1.0).max()
    assert err < 1e-15, err
