import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from physs_gp_b200 import data, likelihood, models, sdes
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
T, n_local, esub = 10000, 32768, 8192
ls_all, steps = bench.make_hypers(n_local, 1)
prior = sdes.BatchedMaternSDE(4, ls_all)
t_host = np.cumsum(steps)
Y_host = torch.empty((n_local, T, 1), dtype=torch.float64, pin_memory=True)
Y_host.copy_(bench.device_observations(n_local, T, dev, seed=1))
out_lml = torch.empty((n_local,), dtype=torch.float64, pin_memory=True)
lik = likelihood.Gaussian(bench.NOISE_VAR)
streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
estarts = list(range(0, n_local, esub))
def step(log):
    for i, s in enumerate(estarts):
        n = min(esub, n_local - s)
        with torch.cuda.stream(streams[i % 3]):
            t0 = time.perf_counter()
            sub_prior = sdes.BatchedMaternSDE(4, prior.ls[s:s + n], prior.var[s:s + n])
            dat = data.TemporalData(t_host, Y_host[s:s + n, :, :, None])
            model = models.SDE_GP(dat, sub_prior, lik)
            t1 = time.perf_counter()
            lml, kf = None, None
            from physs_gp_b200 import filters
            R, R_inv = model._R()
            lml, kf = filters.filter_loop(dat, sub_prior, R=R, R_inv=R_inv, filter_type='b200')
            t2 = time.perf_counter()
            mu, var = filters.smoother_loop(dat, sub_prior, kf, full_state=False, filter_type='b200')
            t3 = time.perf_counter()
            out_lml[s:s + n].copy_(lml, non_blocking=True)
            t4 = time.perf_counter()
            if log: print("sub %d: setup %.1f ms, filter call %.1f ms, smoother call %.1f ms, copy %.1f ms" % (i, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
    torch.cuda.synchronize()
step(False); step(False)
t0 = time.perf_counter(); step(True); print("total %.1f ms" % ((time.perf_counter()-t0)*1e3))
