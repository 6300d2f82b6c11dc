import sys, numpy as np, torch
sys.path.insert(0, '.')
from physs_gp_b200 import ops, sdes
dev = torch.device('cuda:0')
def run(B, T, d, m, L, W):
    rng = np.random.default_rng(0)
    steps = rng.uniform(0.5, 1.5, T) * 0.1
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
    dt_f, dt_s = tt(np.hstack([0.0, steps[1:]])), tt(np.hstack([steps[1:], 0.0]))
    nblk = max(1, d // 4); s = d // nblk
    prior = sdes.BatchedMaternSDE(s, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, nblk))), full_state_obs=(m == d))
    Y = tt(np.sin(0.01 * np.arange(T))[None, :, None] + 0.3 * rng.normal(size=(B, T, m)))
    if B >= 32: Y = Y.transpose(0,1).contiguous().transpose(0,1)
    lam, Pinf = tt(prior.lam()), tt(prior.P_inf())
    H = None if m == d else tt(prior.H())
    disc = ops.Disc.matern(nblk, lam, Pinf)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    R = 0.1 * torch.eye(m, dtype=torch.float64, device=dev)[None, None]
    def timeit(f, n=3):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): out = f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out
    ws = ops.pscan_workspace(B, T, d, L, dev)
    t1, (lml, mf, Pf, st) = timeit(lambda: ops.pscan_filter(dt_f, Y, R, H, m0, Pinf, disc, chunk_len=L, jitter=1e-5, ws=ws, return_status=True))
    t2, (ms, Ps) = timeit(lambda: ops.pscan_smooth(dt_s, mf, Pf, disc, chunk_len=L, jitter=1e-5, ws=ws))
    print(f"B={B} T={T} d={d} m={m} L={L}: scan filter {t1:.2f} ms smoother {t2:.2f} ms status {int(st.item())}")
    for polish in (1, 2):
        t3, o = timeit(lambda: ops.pscan_filter_spec(dt_f, Y, R, H, m0, Pinf, disc, chunk_len=L, warm=W, jitter=1e-5, polish=polish, ws=ws))
        t4, o2 = timeit(lambda: ops.pscan_smooth_spec(dt_s, mf, Pf, disc, chunk_len=L, warm=W, jitter=1e-5, polish=polish, ws=ws))
        print(f"   spec warm={W} polish={polish}: filter {t3:.2f} ms (status {int(o[-1].item())}, relP {float((o[2]-Pf).abs().max()/Pf.abs().max()):.1e} relm {float((o[1]-mf).abs().max()/mf.abs().max()):.1e}) smoother {t4:.2f} ms (status {int(o2[-1].item())}, rel {float((o2[1]-Ps).abs().max()/Ps.abs().max()):.1e})")
run(1, 1000000, 8, 8, 256, 128)
run(1, 1000000, 8, 8, 128, 128)
run(1, 1000000, 4, 1, 64, 64)
run(1, 1000000, 4, 1, 256, 128)
run(1000, 10000, 2, 1, 157, 100)
