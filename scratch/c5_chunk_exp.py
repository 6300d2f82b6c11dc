import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from physs_gp_b200 import ops, sdes
dev = torch.device('cuda:0')
B, T, d = 32768, 10000, 4
ls_all, steps = bench.make_hypers(B, 1)
prior = sdes.BatchedMaternSDE(4, ls_all)
lam = torch.as_tensor(prior.lam(), device=dev); Pinf = torch.as_tensor(prior.P_inf(), device=dev); H = torch.as_tensor(prior.H(), device=dev)
m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
dt_f = torch.as_tensor(np.hstack([0.0, steps[1:]]), device=dev); dt_s = torch.as_tensor(np.hstack([steps[1:], 0.0]), device=dev)
R = torch.full((1, 1, 1, 1), 0.1, dtype=torch.float64, device=dev)
Y = bench.device_observations(B, T, dev, seed=1)
disc = ops.Disc.matern(1, lam, Pinf)
mf = ops.empty_steps(B, T, (d,), dev, True); Pf = ops.empty_steps(B, T, (d, d), dev, True)
ms = ops.empty_steps(B, T, (d,), dev, True); Ps = ops.empty_steps(B, T, (d, d), dev, True)
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
t0, (lml, _, _) = timeit(lambda: ops.kf_filter(dt_f, Y, R, H, m0, Pinf, disc, jitter=1e-5, out=(mf, Pf)))
ref_m = mf[:64].clone(); ref_P = Pf[:64].clone(); ref_l = lml.clone()
print("plain filter %.2f ms" % t0)
t1, _ = timeit(lambda: ops.rts_smooth(dt_s, mf, Pf, disc, jitter=1e-5, out=(ms, Ps)))
ref_ms = ms[:64].clone(); ref_Ps = Ps[:64].clone()
print("plain smoother %.2f ms" % t1)
for L, W, pol in [(5000, 256, 1), (5000, 512, 1), (2500, 256, 1), (3334, 512, 2)]:
    ws = ops.pscan_workspace(B, T, d, L, dev)
    t2, out = timeit(lambda: ops.pscan_filter_spec(dt_f, Y, R, H, m0, Pinf, disc, chunk_len=L, warm=W, jitter=1e-5, polish=pol, out=(mf, Pf), ws=ws))
    print("spec filter L=%d W=%d polish=%d: %.2f ms status %d relm %.1e relP %.1e rellml %.1e" % (L, W, pol, t2, int(out[-1].item()), float((mf[:64]-ref_m).abs().max()/ref_m.abs().max()), float((Pf[:64]-ref_P).abs().max()/ref_P.abs().max()), float(((out[0]-ref_l)/ref_l).abs().max())))
    t3, out2 = timeit(lambda: ops.pscan_smooth_spec(dt_s, mf, Pf, disc, chunk_len=L, warm=W, jitter=1e-5, polish=pol, out=(ms, Ps), ws=ws))
    print("spec smoother: %.2f ms status %d rel %.1e %.1e" % (t3, int(out2[-1].item()), float((ms[:64]-ref_ms).abs().max()/ref_ms.abs().max()), float((Ps[:64]-ref_Ps).abs().max()/ref_Ps.abs().max())))
