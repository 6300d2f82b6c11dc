"""-m gpu: the one-call filter + smoother entry point (physs_kf_filter_smooth_f64) returns exactly what the two
separate entry points return."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["matern", "given"])
@pytest.mark.parametrize("projected", [False, True])
def test_fused_equals_separate(cuda_device, mode, projected):
    from physs_gp_b200 import ops, sdes
    rng = np.random.default_rng(4)
    B, T, s, nblk = 40, 120, 4, 2
    d = s * nblk
    prior = sdes.BatchedMaternSDE(s, rng.uniform(0.5, 1.5, (B, nblk)), rng.uniform(0.5, 1.5, (B, nblk)))
    dev = cuda_device
    tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev)          # noqa: E731
    steps = rng.uniform(0.05, 0.3, T)
    dt_f, dt_s = tt(np.hstack([0.0, steps[1:]])), tt(np.hstack([steps[1:], 0.0]))
    Y = rng.normal(size=(B, T, 1))
    Y[rng.uniform(size=Y.shape) < 0.1] = np.nan
    Yt = tt(Y).transpose(0, 1).contiguous().transpose(0, 1)                       # time-major
    R = torch.full((1, 1, 1, 1), 0.2, dtype=torch.float64, device=dev)
    H, Pinf = tt(prior.H()), tt(prior.P_inf())
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    if mode == "matern":
        disc_f = disc_s = ops.Disc.matern(nblk, tt(prior.lam()), Pinf)
    else:
        def AQ(dts):
            A = np.zeros((B, T, d, d)); Q = np.zeros((B, T, d, d))
            for b in range(B):
                pb = prior.series(b)
                for k, x in enumerate(dts):
                    A[b, k] = pb.expm(None, float(x))
                    Q[b, k] = pb.Q(float(x), A[b, k], prior.P_inf()[b], None)
            return ops.Disc.given(tt(A), tt(Q))
        B = 6
        Yt, Pinf = Yt[:B].contiguous(), Pinf[:B]
        prior = sdes.BatchedMaternSDE(s, prior.ls[:B], prior.var[:B])
        disc_f, disc_s = AQ(dt_f.cpu().numpy()), AQ(dt_s.cpu().numpy())
    Hout = H if projected else None
    lml, mf, Pf = ops.kf_filter(dt_f, Yt, R, H, m0, Pinf, disc_f, jitter=1e-5)
    ms, Ps = ops.rts_smooth(dt_s, mf, Pf, disc_s, Hout=Hout, jitter=1e-5)
    out = ops.kf_filter_smooth(dt_f, dt_s, Yt, R, H, m0, Pinf, disc_f, disc_s, Hout=Hout, jitter=1e-5)
    for a, b in zip(out, (lml, mf, Pf, ms, Ps)):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ packed hand-over (physs_kf_filter_smooth_packed_f64)
def _same_posterior(two_call, packed, bitwise):
    """lml comes from the SAME filter kernel on both paths: always bitwise.  The smoothed outputs come from two
    instantiations of one source (explicit fma, operation for operation): bitwise for full-state outputs on every shape
    measured; with a projected output nvcc makes one different contraction choice in the <4, 2> instantiation, so those
    are held to 1e-13 absolute (values are O(1))."""
    lml, ms, Ps = two_call
    lml2, ms2, Ps2 = packed
    assert torch.equal(lml, lml2)
    assert ms.shape == ms2.shape and Ps.shape == Ps2.shape
    assert bool(torch.isfinite(ms2).all()) and bool(torch.isfinite(Ps2).all())
    if bitwise:
        assert torch.equal(ms, ms2) and torch.equal(Ps, Ps2)
    else:
        assert float((ms - ms2).abs().max()) <= 1e-13 and float((Ps - Ps2).abs().max()) <= 1e-13


def _packed_problem(dev, rng, B, T, s, nblk, mode, m_obs=1):
    from physs_gp_b200 import ops, sdes
    d = s * nblk
    tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev)          # noqa: E731
    steps = rng.uniform(0.05, 0.3, T)
    dt_f, dt_s = tt(np.hstack([0.0, steps[1:]])), tt(np.hstack([steps[1:], 0.0]))
    Y = rng.normal(size=(B, T, m_obs))
    Y[rng.uniform(size=Y.shape) < 0.1] = np.nan
    Yt = tt(Y).transpose(0, 1).contiguous().transpose(0, 1)                       # time-major
    R = tt(0.2 * np.eye(m_obs)).reshape(1, 1, m_obs, m_obs)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    if mode == "iwp":
        H = tt(np.eye(d)[:m_obs][None])
        P0 = tt(np.eye(d)[None] * 2.0)
        disc_f = disc_s = ops.Disc.iwp(tt(rng.uniform(0.5, 1.5, (B, 1))))
        return dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s
    prior = sdes.BatchedMaternSDE(s, rng.uniform(0.5, 1.5, (B, nblk)), rng.uniform(0.5, 1.5, (B, nblk)))
    Hnp = prior.H()
    if m_obs > 1:                                                                # extra rows: other state entries
        Hnp = np.concatenate([Hnp, np.eye(d)[1:m_obs][None].repeat(Hnp.shape[0], 0)], axis=1) if Hnp.ndim == 3 \
            else np.concatenate([Hnp, np.eye(d)[1:m_obs]], axis=0)
    H, Pinf = tt(Hnp), tt(prior.P_inf())
    if mode == "matern":
        disc_f = disc_s = ops.Disc.matern(nblk, tt(prior.lam()), Pinf)
    else:
        def AQ(dts):
            A = np.zeros((B, T, d, d)); Q = np.zeros((B, T, d, d))
            for b in range(B):
                pb = prior.series(b)
                for k, x in enumerate(dts):
                    A[b, k] = pb.expm(None, float(x))
                    Q[b, k] = pb.Q(float(x), A[b, k], prior.P_inf()[b], None)
            return ops.Disc.given(tt(A), tt(Q))                                  # A_k, Q_k are always batch-major
        disc_f, disc_s = AQ(dt_f.cpu().numpy()), AQ(dt_s.cpu().numpy())
    return dt_f, dt_s, Yt, R, H, m0, Pinf, disc_f, disc_s


@pytest.mark.parametrize("shape", [(4, 1), (2, 2), (1, 4), (2, 1), (1, 2), (3, 1), (1, 3), (1, 1)])  # (block size, blocks)
@pytest.mark.parametrize("projected", [False, True])
def test_packed_matches_the_two_output_call_matern(cuda_device, shape, projected):
    """The packed hand-over (14 instead of 20 doubles per step at d = 4) changes no bit of lml and of the full-state
    (ms, Ps); B = 70 leaves a ragged last warp, T = 121 exercises every phase of the 3-stage ring."""
    from physs_gp_b200 import ops
    s, nblk = shape
    args = _packed_problem(cuda_device, np.random.default_rng(11), 70, 121, s, nblk, "matern")
    dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s = args
    assert ops.kf_filter_smooth_packed_supported(Yt, s * nblk, 1, disc_f)
    Hout = H if projected else None
    lml, mf, Pf, ms, Ps = ops.kf_filter_smooth(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=Hout, jitter=1e-5)
    lml2, ms2, Ps2 = ops.kf_filter_smooth_packed(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=Hout, jitter=1e-5)
    _same_posterior((lml, ms, Ps), (lml2, ms2, Ps2), bitwise=not projected)


@pytest.mark.parametrize("case", [("given", 4, 1, 1), ("given", 2, 1, 1), ("given", 3, 1, 1), ("iwp", 4, 1, 1),
                                  ("iwp", 3, 1, 1), ("iwp", 2, 1, 1), ("matern", 4, 1, 2), ("matern", 3, 1, 2),
                                  ("matern", 2, 1, 2)])
def test_packed_other_discretisations_and_m(cuda_device, case):
    """Supplied transitions, integrated-Wiener blocks and two observed outputs per step; T = 2, 3 are the
    prologue / epilogue corner cases of the pipelined smoother."""
    from physs_gp_b200 import ops
    mode, s, nblk, m_obs = case
    B = 33 if mode == "given" else 70
    for T in (1, 2, 3, 40):
        args = _packed_problem(cuda_device, np.random.default_rng(5 + T), B, T, s, nblk, mode, m_obs)
        dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s = args
        if T == 1:
            # one step: [B, 1, m] is contiguous, i.e. batch-major by the layout rule -- the host takes the two calls
            assert not ops.kf_filter_smooth_packed_supported(Yt, s * nblk, m_obs, disc_f)
            continue
        assert ops.kf_filter_smooth_packed_supported(Yt, s * nblk, m_obs, disc_f)
        lml, mf, Pf, ms, Ps = ops.kf_filter_smooth(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H, jitter=1e-5)
        lml2, ms2, Ps2 = ops.kf_filter_smooth_packed(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H,
                                                     jitter=1e-5)
        _same_posterior((lml, ms, Ps), (lml2, ms2, Ps2), bitwise=False)
        full = ops.kf_filter_smooth(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=None, jitter=1e-5)
        full2 = ops.kf_filter_smooth_packed(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=None, jitter=1e-5)
        _same_posterior((full[0], full[3], full[4]), full2, bitwise=True)


def test_packed_rejects_what_it_does_not_cover(cuda_device):
    from physs_gp_b200 import ops
    rng = np.random.default_rng(2)
    dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s = _packed_problem(cuda_device, rng, 40, 30, 4, 2, "matern")
    assert not ops.kf_filter_smooth_packed_supported(Yt, 8, 1, disc_f)           # beyond the register kernels
    with pytest.raises(Exception):
        ops.kf_filter_smooth_packed(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H)
    dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s = _packed_problem(cuda_device, rng, 40, 30, 4, 1, "matern")
    Yb = Yt.contiguous()                                                         # batch-major
    assert not ops.kf_filter_smooth_packed_supported(Yb, 4, 1, disc_f)
    with pytest.raises(NotImplementedError):
        ops.kf_filter_smooth_packed(dt_f, dt_s, Yb, R, H, m0, P0, disc_f, disc_s, Hout=H)
    small = torch.empty(64, dtype=torch.uint8, device=cuda_device)               # a too-small workspace is replaced
    out = ops.kf_filter_smooth_packed(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H, ws=small)
    assert bool(torch.isfinite(out[1]).all())


@pytest.mark.parametrize("order", [4, 3, 2])
def test_model_filter_and_smooth_takes_the_packed_call(cuda_device, order, monkeypatch):
    """SDE_GP.filter_and_smooth(full_state=False) routes a time-major batch through the packed call and returns
    bitwise what the two loops return (settings.fused_packed = False)."""
    from physs_gp_b200 import data, likelihood, models, ops, sdes, settings
    rng = np.random.default_rng(8)
    B, T = 96, 200
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    Y = (np.sin(t)[None] + 0.3 * rng.normal(size=(B, T)))[:, :, None, None]
    Y[rng.uniform(size=Y.shape) < 0.05] = np.nan
    prior = sdes.BatchedMaternSDE(order, rng.uniform(0.6, 1.4, (B, 1)), rng.uniform(0.6, 1.4, (B, 1)))
    model = models.SDE_GP(data.TemporalData(t, Y), prior, likelihood.Gaussian(0.2))
    calls = []
    real = ops.kf_filter_smooth_packed
    monkeypatch.setattr(ops, "kf_filter_smooth_packed", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    monkeypatch.setattr(settings, "fused_packed", True)
    lml, mu, var = model.filter_and_smooth(full_state=False, return_lml=True)
    assert calls == [1]
    monkeypatch.setattr(settings, "fused_packed", False)
    lml2, mu2, var2 = model.filter_and_smooth(full_state=False, return_lml=True)
    assert calls == [1]
    _same_posterior((lml2, mu2, var2), (lml, mu, var), bitwise=False)
