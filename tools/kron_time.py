"""CUDA-event timing of the separable large-block kernels at the config-2 shape (inputs staged once)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from tests.test_gpu_kron import _problem
from physs_gp_b200 import filters, ops, settings

Ns = int(sys.argv[1]) if len(sys.argv) > 1 else 200
T = int(sys.argv[2]) if len(sys.argv) > 2 else 600
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda')
pp, op, t, Y, R = _problem(Ns, T, 1, irregular=False)
dtf = np.concatenate([[0.0], np.diff(t)]); dts = np.concatenate([np.diff(t), [0.0]])
parts = filters._kron_parts(pp, None)
discf, m0, P0 = filters.lower_prior_kron(parts, pp, None, dtf, dev)
discs, _, _ = filters.lower_prior_kron(parts, pp, None, dts, dev)
Yd = torch.as_tensor(Y, device=dev); Rd = torch.as_tensor(R, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(reps):
    torch.cuda.synchronize()
    ev[0].record()
    lml, mf, Pf = ops.kf_filter_kron(Yd, Rd, m0, P0, discf)
    ev[1].record()
    ms, Ps = ops.rts_smooth_kron(mf, Pf, discs, project=True)
    ev[2].record()
    ms2, Ps2 = ops.rts_smooth_kron(mf, Pf, discs, project=False)
    ev[3].record()
    torch.cuda.synchronize()
    print("Ns=%d T=%d: filter %.1f us/step | smoother(projected) %.1f us/step | smoother(full) %.1f us/step | lml %.4f"
          % (Ns, T, 1e3 * ev[0].elapsed_time(ev[1]) / T, 1e3 * ev[1].elapsed_time(ev[2]) / T,
             1e3 * ev[2].elapsed_time(ev[3]) / T, float(lml)), flush=True)
    if ops.kron_prof:
        f = ops.kron_prof["filter"].cpu().numpy() / T / 1e3
        s_ = ops.kron_prof["smoother"].cpu().numpy() / T / 1e3
        print("  filter us/step: F1 %.1f sync %.1f | F2 chol %.1f (gemm %.1f diag %.1f subst %.1f) sync %.1f | F3 trsm %.1f "
              "(load %.1f inv %.1f fwd %.1f bwd %.1f) sync %.1f | F4 %.1f sync %.1f"
              % (f[0], f[1], f[2], f[16], f[17], f[18], f[3], f[4], f[20], f[21], f[22], f[23], f[5], f[6], f[7]))
        print("  smoother rec us/step: B1 gemm %.1f mean %.1f sync %.1f | B2 gemm %.1f sync %.1f"
              % (s_[8], s_[9], s_[10], s_[11], s_[12]))
        pc = s_[32:32 + 2 * 148].reshape(148, 2)
        print("  per-CTA B1 gemm us/step:", np.round(pc[::10, 0], 1), "max %.1f" % pc[:, 0].max())
        print("  per-CTA B2 gemm us/step:", np.round(pc[::10, 1], 1), "max %.1f" % pc[:, 1].max())
        print("  B1 gemm detail (CTA 0) us/step: setup %.2f first-wait %.2f later-waits %.2f compute %.2f epilogue %.2f"
              % (s_[24], s_[25], s_[26], s_[27], s_[28]))
