import sys, numpy as np, torch
sys.path.insert(0, '.')
from physs_gp_b200 import ops, sdes
dev = torch.device('cuda:0')
B, T, d, m, L = 1, 1000000, int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(0)
steps = rng.uniform(0.5, 1.5, T) * 0.1
tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
dt_f, dt_s = tt(np.hstack([0.0, steps[1:]])), tt(np.hstack([steps[1:], 0.0]))
nblk = max(1, d // 4); s = d // nblk
prior = sdes.BatchedMaternSDE(s, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, nblk))), full_state_obs=(m == d))
Y = tt(np.sin(0.01 * np.arange(T))[None, :, None] + 0.3 * rng.normal(size=(B, T, m)))
lam, Pinf = tt(prior.lam()), tt(prior.P_inf())
H = None if m == d else tt(prior.H())
disc = ops.Disc.matern(nblk, lam, Pinf)
m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
R = 0.1 * torch.eye(m, dtype=torch.float64, device=dev)[None, None]
ws = ops.pscan_workspace(B, T, d, L, dev)
for _ in range(2):
    lml, mf, Pf = ops.pscan_filter(dt_f, Y, R, H, m0, Pinf, disc, chunk_len=L, jitter=1e-5, ws=ws)
    ms, Ps = ops.pscan_smooth(dt_s, mf, Pf, disc, chunk_len=L, jitter=1e-5, ws=ws)
torch.cuda.synchronize()
print("ok", float(lml))
