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
