"""-m gpu parity tests of the parallel-in-time path (chunked associative scan, physs_pscan.cu) against the
numpy oracle's SEQUENTIAL filter / smoother (the parity target) and against the CUDA
sequential kernels at sizes the oracle cannot reach.  Tolerance 1e-9 relative (array scale)."""
import numpy as np
import pytest
import torch

from oracle import filters as ofilters
from oracle import sde as osde
from tests import synth
from tests.test_gpu_seq import SPECS, _priors, rel

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _oracle_and_inputs(name, T, seed, jitter, nan_frac=0.08):
    spec, fso = SPECS[name]
    rng = np.random.default_rng(seed)
    pprior, oprior = _priors(spec, fso)
    m = oprior.H().shape[0]
    t = synth.time_grid(T, 0.1, rng)
    Y = synth.noisy_series(1, T, m, rng, nan_frac)[0]
    R = synth.random_spd(rng, (T,), m)
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R, jitter)
    return pprior, oprior, t, Y, R, lml_o, mf_o, Pf_o


@pytest.mark.parametrize("jitter", [1e-5, 0.0])
@pytest.mark.parametrize("name", ["c1_m32", "m52", "m72", "sum_m32x2", "indep_m32x2", "m52_fullstate",
                                  "indep_m32x2_fullstate"])
def test_parallel_filter_smoother_match_sequential_oracle(cuda_device, name, jitter, monkeypatch):
    from physs_gp_b200 import data, filters, settings
    monkeypatch.setattr(settings, "jitter", jitter)
    monkeypatch.setattr(settings, "pscan_chunk_len", 37)          # ragged tail: 500 = 13 * 37 + 19
    T = 500
    pprior, oprior, t, Y, R, lml_o, mf_o, Pf_o = _oracle_and_inputs(name, T, 21, jitter)
    d = data.TemporalData(t, Y[:, :, None])
    lml, kf = filters.filter_loop(d, pprior, R=R, filter_type='b200_parallel')
    assert abs(float(lml) - lml_o) <= TOL * abs(lml_o), (float(lml), lml_o)
    assert rel(kf['m'], mf_o) < TOL and rel(kf['P'], Pf_o) < TOL
    for fs in (True, False):
        ms_o, Ps_o = ofilters.smoother_sequential(oprior, t, mf_o, Pf_o, full_state=fs, jitter=jitter)
        mu, var = filters.smoother_loop(d, pprior, kf, full_state=fs, filter_type='b200_parallel')
        assert rel(mu, ms_o) < TOL and rel(var, Ps_o) < TOL


def _batch_problem(dev, B, T, d, m, given, seed, time_major):
    from physs_gp_b200 import ops, sdes
    rng = np.random.default_rng(seed)
    t = synth.time_grid(T, 0.1, rng)
    dt_f = torch.as_tensor(np.hstack([0.0, np.diff(t)]), device=dev)
    dt_s = torch.as_tensor(np.hstack([np.diff(t), 0.0]), device=dev)
    Y = torch.as_tensor(synth.noisy_series(B, T, m, rng, 0.1), device=dev)
    if time_major:
        Y = Y.transpose(0, 1).contiguous().transpose(0, 1)
    R = torch.as_tensor(synth.random_spd(rng, (B, 1), m), device=dev)
    H = torch.as_tensor(rng.normal(size=(1, m, d)), device=dev) if m != d else None
    if given:
        import scipy.linalg as sla
        Fm = rng.normal(size=(d, d)) * 0.3 - 1.2 * np.eye(d)
        Pinf = sla.solve_continuous_lyapunov(Fm, -np.eye(d))
        uniq, inv = np.unique(np.hstack([0.0, np.diff(t)]), return_inverse=True)
        Au = np.stack([sla.expm(Fm * x) for x in uniq])
        A = Au[inv]
        Q = Pinf - A @ Pinf @ np.swapaxes(A, -1, -2)
        A_s = np.concatenate([A[1:], np.eye(d)[None]])
        Q_s = np.concatenate([Q[1:], np.zeros((1, d, d))])
        disc_f = ops.Disc.given(torch.as_tensor(A[None], device=dev), torch.as_tensor(Q[None], device=dev))
        disc_s = ops.Disc.given(torch.as_tensor(A_s[None], device=dev), torch.as_tensor(Q_s[None], device=dev))
        P0 = torch.as_tensor(Pinf[None], device=dev)
    else:
        s_blk = d if d <= 4 else (4 if d % 4 == 0 else 3)
        prior = sdes.BatchedMaternSDE(s_blk, synth.log_uniform(rng, 0.5, 2.0, (B, d // s_blk)))
        lam = torch.as_tensor(prior.lam(), device=dev)
        P0 = torch.as_tensor(prior.P_inf(), device=dev)
        disc_f = disc_s = ops.Disc.matern(d // s_blk, lam, P0)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    return dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s


@pytest.mark.parametrize("B,T,d,m,given,time_major,chunk", [
    (3, 20000, 4, 1, False, False, 256), (70, 3000, 4, 1, False, True, 200), (2, 8000, 6, 6, False, False, 128),
    (2, 8000, 8, 3, True, False, 100), (1, 30000, 12, 12, False, False, 64), (1, 4000, 24, 2, False, False, 250),
    (5, 1000, 3, 1, True, False, 1000), (4, 999, 2, 2, False, False, 1)])
def test_parallel_equals_sequential_cuda(cuda_device, B, T, d, m, given, time_major, chunk):
    """Long series / odd shapes: parallel-in-time == sequential CUDA path (itself oracle-checked)."""
    from physs_gp_b200 import ops
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(cuda_device, B, T, d, m, given, 7 + d, time_major)
    # one-step chunks cannot host a fix-up pass (nothing to contract over): exact-scan case only
    for jitter in ((1e-5, 0.0) if chunk >= 16 else (0.0,)):
        lml, mf, Pf = ops.kf_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter=jitter)
        ms, Ps = ops.rts_smooth(dt_s, mf, Pf, disc_s, jitter=jitter)
        lml2, mf2, Pf2, st = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=chunk, jitter=jitter,
                                              patience=min(4, chunk), return_status=True)
        ms2, Ps2 = ops.pscan_smooth(dt_s, mf, Pf, disc_s, chunk_len=chunk, jitter=jitter)
        torch.cuda.synchronize()
        assert int(st.item()) == 0
        assert rel(lml2, lml.cpu().numpy()) < TOL
        assert rel(mf2, mf.cpu().numpy()) < TOL and rel(Pf2, Pf.cpu().numpy()) < TOL
        assert rel(ms2, ms.cpu().numpy()) < TOL and rel(Ps2, Ps.cpu().numpy()) < TOL


def test_scan_alone_is_exact_without_jitter(cuda_device):
    """jitter = 0, no polish pass: the boundary states come from the associative scan only."""
    from physs_gp_b200 import ops
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(cuda_device, 2, 5000, 8, 2, False, 3, False)
    lml, mf, Pf = ops.kf_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter=0.0)
    lml2, mf2, Pf2 = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=50, jitter=0.0, polish=0)
    assert rel(mf2, mf.cpu().numpy()) < TOL and rel(Pf2, Pf.cpu().numpy()) < TOL and rel(lml2, lml.cpu().numpy()) < TOL


def test_unconverged_flag_is_raised(cuda_device):
    """A chunk too short for the fix-up to settle (patience longer than the chunk) must raise the flag."""
    from physs_gp_b200 import ops
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(cuda_device, 1, 400, 4, 1, False, 5, False)
    out = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=8, jitter=1e-5, polish=1, patience=50,
                           return_status=True)
    assert int(out[-1].item()) == 1


@pytest.mark.parametrize("d,m,given", [(4, 1, False), (8, 8, False), (6, 2, True)])
def test_time_shards_on_one_gpu(cuda_device, d, m, given):
    """The multi-GPU building blocks (local -> [all-gather] -> fold -> finish) run for 3 time ranges one
    after the other on one GPU; the concatenation must equal the sequential result of the whole series."""
    from physs_gp_b200 import ops, timeshard
    B, T, chunk = 3, 2900, 64
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(cuda_device, B, T, d, m, given, 31 + d, False)
    jitter = 0.0
    lml, mf, Pf = ops.kf_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter=jitter)
    ms, Ps = ops.rts_smooth(dt_s, mf, Pf, disc_s, jitter=jitter)
    ranges = timeshard.time_ranges(T, 3)

    def sl_disc(disc, t0, t1):
        if disc.mode == 0:
            return ops.Disc.given(disc.A[:, t0:t1].contiguous(), disc.Q[:, t0:t1].contiguous())
        return disc
    Rb = R.expand(B, T, m, m)
    loc = []
    for (t0, t1) in ranges:
        ws = ops.pscan_workspace(B, t1 - t0, d, chunk, cuda_device)
        args = (dt_f[t0:t1], Y[:, t0:t1].contiguous(), Rb[:, t0:t1].contiguous(), H, m0, P0, sl_disc(disc_f, t0, t1))
        loc.append((ws, args, ops.pscan_filter_local(*args, chunk, ws, jitter=jitter)))
    totals = torch.stack([x[2] for x in loc])
    lml_sum, mfs, Pfs = 0.0, [], []
    for r, (ws, args, _) in enumerate(loc):
        start = ops.pscan_filter_fold(totals[:r], m0.expand(B, d), P0.expand(B, d, d)) if r > 0 else None
        l, a, b_, st = ops.pscan_filter_finish(*args, chunk, ws, start=start, jitter=jitter)
        lml_sum = lml_sum + l
        mfs.append(a), Pfs.append(b_)
    mf2, Pf2 = torch.cat(mfs, 1), torch.cat(Pfs, 1)
    assert rel(lml_sum, lml.cpu().numpy()) < TOL
    assert rel(mf2, mf.cpu().numpy()) < TOL and rel(Pf2, Pf.cpu().numpy()) < TOL
    # smoother, backwards
    sloc = []
    for r, (t0, t1) in enumerate(ranges):
        ws = loc[r][0]
        sargs = (dt_s[t0:t1], mfs[r], Pfs[r], sl_disc(disc_s, t0, t1))
        sloc.append((ws, sargs, ops.pscan_smooth_local(*sargs, chunk, ws, jitter=jitter)))
    stotals = torch.stack([x[2] for x in sloc])
    mss, Pss = [], []
    for r, (ws, sargs, _) in enumerate(sloc):
        start = None
        if r < len(ranges) - 1:
            start = ops.pscan_smooth_fold(stotals[r + 1:], mfs[-1][:, -1].contiguous(), Pfs[-1][:, -1].contiguous())
        a, b_ = ops.pscan_smooth_finish(*sargs, chunk, ws, start=start, jitter=jitter)
        mss.append(a), Pss.append(b_)
    assert rel(torch.cat(mss, 1), ms.cpu().numpy()) < TOL and rel(torch.cat(Pss, 1), Ps.cpu().numpy()) < TOL


@pytest.mark.parametrize("B,T,d,m,given,time_major,chunk,warm", [
    (3, 20000, 4, 1, False, False, 256, 128), (70, 3000, 4, 1, False, True, 200, 100),
    (2, 8000, 8, 8, False, False, 128, 128), (2, 8000, 8, 3, True, False, 100, 100),
    (1, 9000, 12, 2, False, False, 250, 200), (40, 2500, 2, 1, False, True, 97, 97)])
def test_speculative_parallel_contract(cuda_device, B, T, d, m, given, time_major, chunk, warm):
    """Speculative mode (warm-up + verify / repair passes instead of summaries + scan).  Contract: when the
    device flag is 0 the result equals the sequential kernels' to 1e-9; when it is 1 (slowly mixing filters:
    the last case forgets a wrong start only at 0.92 per step) the caller must use the exact scan, which is
    checked to hold parity on the same inputs."""
    from physs_gp_b200 import ops
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(cuda_device, B, T, d, m, given, 17 + d, time_major)
    jitter = 1e-5
    lml, mf, Pf = ops.kf_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter=jitter)
    ms, Ps = ops.rts_smooth(dt_s, mf, Pf, disc_s, jitter=jitter)
    lml2, mf2, Pf2, st = ops.pscan_filter_spec(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=chunk, warm=warm, jitter=jitter,
                                               polish=6)
    ms2, Ps2, st2 = ops.pscan_smooth_spec(dt_s, mf, Pf, disc_s, chunk_len=chunk, warm=warm, jitter=jitter, polish=6)
    torch.cuda.synchronize()
    if int(st.item()) == 0:
        assert rel(lml2, lml.cpu().numpy()) < TOL
        assert rel(mf2, mf.cpu().numpy()) < TOL and rel(Pf2, Pf.cpu().numpy()) < TOL
    else:
        # slow mixing: the O(jitter) boundary error of the scan also needs more fix-up passes than the default 4
        lml2, mf2, Pf2, st = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=chunk, jitter=jitter,
                                              polish=8, return_status=True)
        assert int(st.item()) == 0
        assert rel(mf2, mf.cpu().numpy()) < TOL and rel(Pf2, Pf.cpu().numpy()) < TOL
    if int(st2.item()) == 0:
        assert rel(ms2, ms.cpu().numpy()) < TOL and rel(Ps2, Ps.cpu().numpy()) < TOL
    else:
        ms2, Ps2 = ops.pscan_smooth(dt_s, mf, Pf, disc_s, chunk_len=chunk, jitter=jitter)
        assert rel(ms2, ms.cpu().numpy()) < TOL and rel(Ps2, Ps.cpu().numpy()) < TOL
    if d >= 8 or (d == 4 and B == 70):
        assert int(st.item()) == 0 and int(st2.item()) == 0      # the fast-mixing cases must take the fast path


def test_speculative_mode_flags_insufficient_warmup(cuda_device):
    """A warm-up far shorter than the filter's memory, with a single verification pass that cannot repair a
    whole chunk within `patience`-limited agreement, must raise the flag (the caller then uses the exact scan)."""
    from physs_gp_b200 import ops
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(cuda_device, 1, 600, 4, 1, False, 5, False)
    out = ops.pscan_filter_spec(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=8, warm=1, jitter=1e-5, polish=1, patience=50)
    assert int(out[-1].item()) == 1


def test_host_api_recovers_from_unconverged_scan(cuda_device, monkeypatch):
    """filter_type='b200_parallel' with chunks far shorter than the filter's memory and a single fix-up pass: the
    host API must still return the SEQUENTIAL result at 1e-9 -- by retrying with more passes, or, if that is not
    enough either, through the sequential kernels (physs_gp_b200/filters.py:_filter_impl)."""
    import warnings
    from physs_gp_b200 import data, filters, kernels as K, sdes, settings
    from oracle import filters as ofilters
    from oracle import sde as osde
    rng = np.random.default_rng(8)
    T = 1500
    t = np.cumsum(rng.uniform(0.5, 1.5, T) * 0.02)
    Y = np.sin(0.01 * np.arange(T))[:, None] + 0.3 * rng.normal(size=(T, 1))
    R = np.tile(0.5 * np.eye(1), [T, 1, 1])
    prior = sdes.LTI_SDE(sdes.Independent([K.Matern72(3.0, 1.0)]))       # lengthscale = 150 steps
    oprior = osde.LTI_SDE([osde.Matern72(3.0, 1.0)])
    monkeypatch.setattr(settings, "pscan_chunk_len", 16)
    monkeypatch.setattr(settings, "pscan_polish", 1)
    d = data.TemporalData(t, Y[:, :, None])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lml, kf = filters.filter_loop(d, prior, R=R, filter_type="b200_parallel")
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(oprior, t, Y, R, 1e-5)
    assert abs(float(lml) - lml_o) <= TOL * abs(lml_o)
    assert rel(kf['m'], mf_o) < TOL and rel(kf['P'], Pf_o) < TOL


@pytest.mark.parametrize("B,T,d,m,time_major", [(40, 3000, 2, 1, True), (3, 4000, 4, 1, False), (2, 2500, 8, 8, False)])
def test_fixup_early_out_is_bitwise_neutral(cuda_device, monkeypatch, B, T, d, m, time_major):
    """Fix-up passes after a pass in which no recomputed step disagreed are skipped on the device (pass_changed /
    prev_changed flags); the result must be bitwise the one of running all passes (PHYSS_PSCAN_NO_EARLY_OUT=1), for the
    register kernels (d <= 4) and the register-tile kernels (d = 8), with jitter (the case that needs the passes)."""
    from physs_gp_b200 import ops
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(cuda_device, B, T, d, m, False, 40 + d, time_major)
    outs = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("PHYSS_PSCAN_NO_EARLY_OUT", "1")
        else:
            monkeypatch.delenv("PHYSS_PSCAN_NO_EARLY_OUT", raising=False)
        lml, mf, Pf, st = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=100, jitter=1e-5, polish=6,
                                           return_status=True)
        torch.cuda.synchronize()
        assert int(st.item()) == 0
        outs.append((lml.clone(), mf.clone(), Pf.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_sum_steps_matches_torch(cuda_device):
    """physs_sum_steps_f64 (the ELBO's per-series ELL sums) in both step layouts, with and without the subtracted
    array, ragged sizes."""
    from physs_gp_b200 import cvi
    g = torch.Generator(device=cuda_device).manual_seed(5)
    for B, T in ((1000, 3001), (37, 513), (1, 20000), (70, 1)):
        x = torch.randn((B, T), dtype=torch.float64, device=cuda_device, generator=g)
        y = torch.randn((B, T), dtype=torch.float64, device=cuda_device, generator=g)
        for tm in (False, True):
            xx = x.t().contiguous().t() if tm else x
            yy = y.t().contiguous().t() if tm else y
            ref = x.sum(-1)
            got = cvi.sum_steps(xx)
            assert float((got - ref).abs().max()) <= 1e-12 * max(1.0, float(ref.abs().max()))
            got2 = cvi.sum_steps(xx, yy)
            ref2 = (x - y).sum(-1)
            assert float((got2 - ref2).abs().max()) <= 1e-12 * max(1.0, float(ref2.abs().max()))
