"""Worker for tests/test_timeshard_host.py: world_size ranks over gloo on CPU run
physs_gp_b200.timeshard.filter_smooth with the oracle-backed ops stand-in and write their local results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import sde as osde  # noqa: E402
from physs_gp_b200 import timeshard  # noqa: E402
from tests.dist import oracle_ops  # noqa: E402


def problem(seed=3, B=2, T=41):
    rng = np.random.default_rng(seed)
    prior = osde.LTI_SDE([osde.Matern52(0.8, 1.2)])
    t = np.cumsum(rng.uniform(0.5, 1.5, T) * 0.1)
    Y = np.sin(0.3 * np.arange(T))[None, :, None] + 0.3 * rng.normal(size=(B, T, 1))
    Y[rng.uniform(size=Y.shape) < 0.15] = np.nan
    R = np.tile(0.1 * np.eye(1), [B, T, 1, 1])
    dt_f = np.hstack([0.0, np.diff(t)])
    dt_s = np.hstack([np.diff(t), 0.0])
    Pinf = prior.P_inf()
    A_f = np.stack([prior.expm(x) for x in dt_f]); Q_f = np.stack([prior.Q(x, a, Pinf) for x, a in zip(dt_f, A_f)])
    A_s = np.stack([prior.expm(x) for x in dt_s]); Q_s = np.stack([prior.Q(x, a, Pinf) for x, a in zip(dt_s, A_s)])
    return prior, t, Y, R, dt_f, dt_s, A_f, Q_f, A_s, Q_s


def main():
    out_dir = sys.argv[1]
    dist.init_process_group("gloo")
    comm = timeshard.TorchDist()
    prior, t, Y, R, dt_f, dt_s, A_f, Q_f, A_s, Q_s = problem()
    T = len(t)
    t0, t1 = timeshard.time_ranges(T, comm.world)[comm.rank]
    sl = slice(t0, t1)
    tt = torch.as_tensor
    res = timeshard.filter_smooth(
        comm, oracle_ops, tt(dt_f[sl]), tt(dt_s[sl]), tt(Y[:, sl]), tt(R[:, sl]), tt(prior.H()[None]),
        tt(prior.m_inf()[:, 0][None]), tt(prior.P_inf()[None]), oracle_ops.Disc(A_f[sl], Q_f[sl]),
        oracle_ops.Disc(A_s[sl], Q_s[sl]), chunk_len=8, jitter=0.0)
    lml, mf, Pf, ms, Ps, status = res
    np.savez(os.path.join(out_dir, "rank%d.npz" % comm.rank), lml=lml.numpy(), mf=mf.numpy(), Pf=Pf.numpy(),
             ms=ms.numpy(), Ps=Ps.numpy(), t0=t0, t1=t1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
