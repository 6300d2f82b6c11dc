"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_gpu_pscan import _batch_problem
from physs_gp_b200 import ops, cvi, sdes
dev = torch.device('cuda:0')
for (B, T, d, m, given, tm) in [(70, 40, 4, 1, False, True), (5, 40, 3, 3, True, False), (3, 60, 8, 8, False, False),
                                (2, 50, 12, 2, True, False), (2, 40, 20, 1, False, False), (33, 40, 8, 1, False, True)]:
    dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(dev, B, T, d, m, given, 1, tm)
    lml, mf, Pf = ops.kf_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter=1e-5)
    ms, Ps = ops.rts_smooth(dt_s, mf, Pf, disc_s, jitter=1e-5)
    lml2, mf2, Pf2, st = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=16, jitter=1e-5, return_status=True)
    ms2, Ps2 = ops.pscan_smooth(dt_s, mf, Pf, disc_s, chunk_len=16, jitter=1e-5)
    torch.cuda.synchronize()
    print(B, T, d, m, float((mf2 - mf).abs().max()), float((Ps2 - Ps).abs().max()), int(st.item()))
rng = np.random.default_rng(0)
for D, P in [(2, 1), (4, 4), (6, 2)]:
    N = 50
    G = rng.normal(size=(N, D, D)) * 0.3
    V = torch.as_tensor(G @ np.swapaxes(G, -1, -2) + 0.5 * np.eye(D), device=dev)
    Yt = torch.as_tensor(rng.normal(size=(N, D)), device=dev)
    y = torch.as_tensor(rng.integers(0, 4, size=(N, P)).astype(float), device=dev)
    W = None if P == D else torch.as_tensor(rng.normal(size=(P, D)), device=dev)
    out = cvi.natgrad_step(Yt, V, Yt * 0.5, V * 0.1, y, W, cvi.PoissonLik(1.0), 0.3, want_ell=True)
    torch.cuda.synchronize()
    print("cvi", D, P, float(out[2].sum()))
print("sanit ok")
