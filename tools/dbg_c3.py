"""c3 debug: which T / DMMA setting breaks the parallel-in-time pass (status flag, lml finiteness)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from physs_gp_b200 import ops, sdes, timeshard
dev = torch.device("cuda", 0)
d = m = 8
for T in (1000000, 2000000, 4000000, 8000000):
    rng = np.random.default_rng(0)
    steps = rng.uniform(0.5, 1.5, T) * 0.1
    dt_f, dt_s = np.hstack([0.0, steps[1:]]), np.hstack([steps[1:], 0.0])
    prior = sdes.BatchedMaternSDE(4, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (1, 2))) * 1.0, full_state_obs=True)
    L = ops.even_chunk_len(T, 256)
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
    Y = tt(np.sin(0.01 * np.arange(T))[None, :, None] + 0.3 * rng.normal(size=(1, T, m)))
    lam, Pinf = tt(prior.lam()), tt(prior.P_inf())
    disc = ops.Disc.matern(2, lam, Pinf)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    R = 0.1 * torch.eye(m, dtype=torch.float64, device=dev)[None, None]
    ws = ops.pscan_workspace(1, T, d, L, dev)
    out = timeshard.filter_smooth(timeshard.SingleProcess(), ops, tt(dt_f), tt(dt_s), Y, R, None, m0, Pinf, disc, disc,
                                  chunk_len=L, jitter=1e-5, ws=ws)
    torch.cuda.synchronize()
    lml, mf, Pf, ms, Ps, st = out
    print(T, L, "lml", float(lml[0]), "status", int(st.item()), "finite", bool(torch.isfinite(ms).all()), bool(torch.isfinite(Ps).all()), flush=True)
