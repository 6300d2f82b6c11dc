import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from physs_gp_b200 import data, likelihood, models, sdes, filters
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
T = 10000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ls_all, steps = bench.make_hypers(n, 1)
prior = sdes.BatchedMaternSDE(4, ls_all)
t_host = np.cumsum(steps)
Yd = bench.device_observations(n, T, dev, seed=1)        # [n, T, 1] device
print("Y layout time-major:", Yd.transpose(0,1).is_contiguous(), "contiguous:", Yd.is_contiguous())
Yb = Yd.contiguous()
lik = likelihood.Gaussian(bench.NOISE_VAR)
def run(Y, full_state):
    dat = data.TemporalData(t_host, Y[:, :, :, None])
    model = models.SDE_GP(dat, prior, lik)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    R, R_inv = model._R()
    lml, kf = filters.filter_loop(dat, prior, R=R, R_inv=R_inv, filter_type='b200')
    e[1].record()
    mu, var = filters.smoother_loop(dat, prior, kf, full_state=full_state, filter_type='b200')
    e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
for name, Y in (("batch-major Y (as H2D delivers)", Yb), ("time-major Y", Yd)):
    for fs in (False, True):
        run(Y, fs); run(Y, fs)
        t0 = time.perf_counter(); f, s = run(Y, fs); w = (time.perf_counter() - t0) * 1e3
        print("%s full_state=%s: filter %.1f ms, smoother %.1f ms, wall %.1f ms" % (name, fs, f, s, w))
