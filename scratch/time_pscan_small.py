import sys, numpy as np, torch
sys.path.insert(0, '.')
from physs_gp_b200 import ops, sdes
dev = torch.device('cuda:0')
def run(B, T, s, chunks, jitter=1e-5):
    rng = np.random.default_rng(0)
    steps = rng.uniform(0.5, 1.5, T) * 0.1
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
    dt_f, dt_s = tt(np.hstack([0.0, steps[1:]])), tt(np.hstack([steps[1:], 0.0]))
    prior = sdes.BatchedMaternSDE(s, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1))))
    Y = tt(np.sin(0.01 * np.arange(T))[None, :, None] + 0.3 * rng.normal(size=(B, T, 1)))
    Y = Y.transpose(0,1).contiguous().transpose(0,1)
    lam, Pinf, H = tt(prior.lam()), tt(prior.P_inf()), tt(prior.H())
    disc = ops.Disc.matern(1, lam, Pinf)
    m0 = torch.zeros((1, s), dtype=torch.float64, device=dev)
    R = 0.1 * torch.ones((1,1,1,1), dtype=torch.float64, device=dev)
    def timeit(f, n=3):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): out = f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out
    tf, (lml, mf, Pf) = timeit(lambda: ops.kf_filter(dt_f, Y, R, H, m0, Pinf, disc, jitter=jitter))
    ts, _ = timeit(lambda: ops.rts_smooth(dt_s, mf, Pf, disc, jitter=jitter))
    print(f"B={B} T={T} d={s} sequential: filter {tf:.2f} ms smoother {ts:.2f} ms")
    for L in chunks:
        ws = ops.pscan_workspace(B, T, s, L, dev)
        mfo, Pfo = torch.empty_like(mf), torch.empty_like(Pf)
        for polish in (0, 4):
            tf2, out = timeit(lambda: ops.pscan_filter(dt_f, Y, R, H, m0, Pinf, disc, chunk_len=L, jitter=jitter, polish=polish, ws=ws, out=(mfo, Pfo), return_status=True))
            print(f"  chunk {L:5d} polish {polish}: pscan filter {tf2:.2f} ms status {int(out[-1].item())} relP {float((out[2]-Pf).abs().max()/Pf.abs().max()):.1e} relm {float((out[1]-mf).abs().max()/mf.abs().max()):.1e}")
        mso, Pso = torch.empty_like(mf), torch.empty_like(Pf)
        ts2, _ = timeit(lambda: ops.pscan_smooth(dt_s, mf, Pf, disc, chunk_len=L, jitter=jitter, ws=ws, out=(mso, Pso)))
        print(f"  chunk {L:5d}          : pscan smoother {ts2:.2f} ms")
run(1000, 10000, 2, [50, 100, 157, 400])
run(1, 1000000, 4, [64, 256, 1024])
run(1, 10000, 2, [32, 64])
