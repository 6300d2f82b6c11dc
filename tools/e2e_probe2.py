import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from physs_gp_b200 import data, likelihood, models, sdes, filters
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
T, n_local = 10000, 65536
esub = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nstreams = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ls_all, steps = bench.make_hypers(n_local, 1)
prior = sdes.BatchedMaternSDE(4, ls_all)
t_host = np.cumsum(steps)
Y_host = torch.empty((n_local, T, 1), dtype=torch.float64, pin_memory=True)
for s0 in range(0, n_local, 16384):
    Y_host[s0:s0+16384].copy_(bench.device_observations(16384, T, dev, seed=1 + s0))
out_mu = torch.empty((n_local, T, 1), dtype=torch.float64, pin_memory=True)
out_var = torch.empty((n_local, T, 1), dtype=torch.float64, pin_memory=True)
out_lml = torch.empty((n_local,), dtype=torch.float64, pin_memory=True)
lik = likelihood.Gaussian(bench.NOISE_VAR)
streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
estarts = list(range(0, n_local, esub))
def step(log):
    evs = []
    base = torch.cuda.Event(enable_timing=True); base.record()
    for i, s in enumerate(estarts):
        n = min(esub, n_local - s)
        with torch.cuda.stream(streams[i % nstreams]):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            sub_prior = sdes.BatchedMaternSDE(4, prior.ls[s:s + n], prior.var[s:s + n])
            dat = data.TemporalData(t_host, Y_host[s:s + n, :, :, None])
            model = models.SDE_GP(dat, sub_prior, lik)
            R, R_inv = model._R()
            lml, kf = filters.filter_loop(dat, sub_prior, R=R, R_inv=R_inv, filter_type='b200')
            e[1].record()
            mu, var = filters.smoother_loop(dat, sub_prior, kf, full_state=False, filter_type='b200')
            e[2].record()
            out_mu[s:s + n].copy_(mu[..., 0], non_blocking=True)
            out_var[s:s + n].copy_(var[..., 0], non_blocking=True)
            out_lml[s:s + n].copy_(lml, non_blocking=True)
            e[3].record()
            evs.append(e)
    torch.cuda.synchronize()
    if log:
        for i, e in enumerate(evs):
            print("sub %d: start %.0f  filter(+H2D) done %.0f  smoother done %.0f  D2H done %.0f" % (i, base.elapsed_time(e[0]), base.elapsed_time(e[1]), base.elapsed_time(e[2]), base.elapsed_time(e[3])))
step(False); step(False)
t0 = time.perf_counter(); step(True); el = time.perf_counter()-t0; print("total %.1f ms -> %.3g state-steps/s" % (el*1e3, n_local*T/el))
