"""Driver for the ncu capture of the posterior-only (packed hand-over) call at the bench sub-batch:
32,768 Matern-7/2 series x 600 steps, one warm-up call, one profiled call (tools/prof_packed.sh)."""
import numpy as np
import torch

from physs_gp_b200 import ops, sdes

B, T = 32768, 600
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
prior = sdes.BatchedMaternSDE(4, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1))))
tt = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev)          # noqa: E731
steps = rng.uniform(0.05, 0.15, T)
dt_f, dt_s = tt(np.hstack([0.0, steps[1:]])), tt(np.hstack([steps[1:], 0.0]))
Y = torch.randn((T, B, 1), dtype=torch.float64, device=dev).transpose(0, 1)
R = torch.full((1, 1, 1, 1), 0.1, dtype=torch.float64, device=dev)
H, Pinf, lam = tt(prior.H()), tt(prior.P_inf()), tt(prior.lam())
m0 = torch.zeros((1, 4), dtype=torch.float64, device=dev)
disc = ops.Disc.matern(1, lam, Pinf)
for _ in range(2):
    lml, ms, Ps = ops.kf_filter_smooth_packed(dt_f, dt_s, Y, R, H, m0, Pinf, disc, Hout=H, jitter=1e-5)
    torch.cuda.synchronize()
assert bool(torch.isfinite(lml).all())
print("ok", float(lml.mean()))
