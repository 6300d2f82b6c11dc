"""Multi-GPU check of the time-sharded path over NCCL (run on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        tests/dist/run_timeshard_nccl.py [T] [d]

Every rank filters + smooths its own time range of B long series (physs_gp_b200.timeshard.filter_smooth,
all-gather of the range summaries over NCCL); rank 0 then also runs the whole series sequentially on its
own GPU and checks every rank's range against it (1e-9 relative).  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from physs_gp_b200 import ops, sdes, timeshard  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = timeshard.TorchDist()
    B, m, chunk, jitter = 2, 1, 256, 1e-5
    rng = np.random.default_rng(0)
    steps = rng.uniform(0.5, 1.5, T) * 0.1
    dt_f = np.hstack([0.0, steps[1:]])
    dt_s = np.hstack([steps[1:], 0.0])
    nblk = d // 4
    prior = sdes.BatchedMaternSDE(4, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, nblk))))
    Y = np.sin(0.01 * np.arange(T))[None, :, None] + 0.3 * rng.normal(size=(B, T, m))
    Y[rng.uniform(size=Y.shape) < 0.05] = np.nan
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    lam, Pinf, H = tt(prior.lam()), tt(prior.P_inf()), tt(prior.H())
    disc = ops.Disc.matern(nblk, lam, Pinf)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    R = torch.full((1, 1, m, m), 0.1, dtype=torch.float64, device=dev)
    t0, t1 = timeshard.time_ranges(T, comm.world)[comm.rank]
    args = (tt(dt_f[t0:t1]), tt(dt_s[t0:t1]), tt(Y[:, t0:t1]), R, H, m0, Pinf, disc, disc)
    for _ in range(2):                                   # warm-up + timed
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lml, mf, Pf, ms, Ps, status = timeshard.filter_smooth(comm, ops, *args, chunk_len=chunk, jitter=jitter,
                                                              cross_rank_polish=True)
        e1.record()
        torch.cuda.synchronize()
    ms_t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    # reference: whole series, sequential kernels, on this rank's GPU
    lml_r, mf_r, Pf_r = ops.kf_filter(tt(dt_f), tt(Y), R, H, m0, Pinf, disc, jitter=jitter)
    ms_r, Ps_r = ops.rts_smooth(tt(dt_s), mf_r, Pf_r, disc, jitter=jitter)

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    errs = torch.tensor([rel(lml, lml_r), rel(mf, mf_r[:, t0:t1]), rel(Pf, Pf_r[:, t0:t1]), rel(ms, ms_r[:, t0:t1]),
                         rel(Ps, Ps_r[:, t0:t1]), float(status.item())], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if comm.rank == 0:
        e = errs.tolist()
        line = {"test": "time-sharded filter+smoother over NCCL", "n_gpus": comm.world, "B": B, "T": T, "d": d,
                "chunk_len": chunk, "jitter": jitter, "ms": float(ms_t.item()),
                "state_steps_per_s": B * T / (float(ms_t.item()) * 1e-3),
                "max_rel_err": {"lml": e[0], "mf": e[1], "Pf": e[2], "ms": e[3], "Ps": e[4]},
                "unconverged": e[5], "ok": bool(max(e[:5]) < 1e-9)}
        print(json.dumps(line), flush=True)
    ok = max(errs.tolist()[:5]) < 1e-9
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
