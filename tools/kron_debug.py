"""Diagnostic run of the separable large-block kernels: per-quantity, per-step errors vs the numpy oracle (no asserts)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import filters as ofilters
from tests.test_gpu_kron import _problem
from tests.test_gpu_seq import rel
from physs_gp_b200 import data, filters, settings

def run(Ns, T, kind="m32", jitter=1e-5, dense_R=False, nan_frac=0.05):
    settings.jitter = jitter
    pp, op, t, Y, R = _problem(Ns, T, 3 + Ns, kind=kind, dense_R=dense_R, nan_frac=nan_frac)
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(op, t, Y, R, jitter)
    d = data.TemporalData(t, Y[:, :, None])
    try:
        lml, kf = filters.filter_loop(d, pp, R=R)
        torch.cuda.synchronize()
    except Exception as e:
        print("FILTER FAILED", Ns, T, kind, repr(e)); return
    errP = [rel(kf['P'][k], Pf_o[k]) for k in range(T)]
    errm = [rel(kf['m'][k], mf_o[k]) for k in range(T)]
    print("Ns=%d T=%d %s jit=%g | filter P err first %s max %.2e | m err first %s max %.2e | lml %.12g vs %.12g"
          % (Ns, T, kind, jitter, ["%.1e" % e for e in errP[:3]], max(errP), ["%.1e" % e for e in errm[:3]], max(errm),
             float(lml), lml_o))
    for fs in (True, False):
        ms_o, Ps_o = ofilters.smoother_sequential(op, t, mf_o, Pf_o, full_state=fs, jitter=jitter)
        try:
            kf_o = {'m': torch.as_tensor(mf_o, device='cuda'), 'P': torch.as_tensor(Pf_o, device='cuda')}
            mu, var = filters.smoother_loop(d, pp, kf_o, full_state=fs)
            torch.cuda.synchronize()
        except Exception as e:
            print("SMOOTHER FAILED", repr(e)); return
        eP = [rel(var[k], Ps_o[k]) for k in range(T)]
        em = [rel(mu[k], ms_o[k]) for k in range(T)]
        print("   smoother fs=%s (oracle filter in): P err last3 %s max %.2e | m err last3 %s max %.2e"
              % (fs, ["%.1e" % e for e in eP[-3:]], max(eP), ["%.1e" % e for e in em[-3:]], max(em)))

if __name__ == "__main__":
    run(20, 6); run(20, 6, jitter=0.0); run(40, 5); run(37, 5); run(70, 4, dense_R=True)
    run(15, 5, kind="m52"); run(12, 5, kind="m72")
    run(17, 700)
    # timing at the config-2 shape
    Ns, T = 200, 300
    pp, op, t, Y, R = _problem(Ns, T, 1, dense_R=False)
    d = data.TemporalData(t, Y[:, :, None])
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lml, kf = filters.filter_loop(d, pp, R=R); torch.cuda.synchronize(); t1 = time.perf_counter()
        mu, var = filters.smoother_loop(d, pp, kf, full_state=False); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("Ns=200 T=%d: filter %.1f us/step, smoother %.1f us/step, lml %.6f finite=%s"
              % (T, 1e6 * (t1 - t0) / T, 1e6 * (t2 - t1) / T, float(lml), bool(torch.isfinite(var).all())))
    settings.kron_kernels = False
    lml2, kf2 = filters.filter_loop(d, pp, R=R); mu2, var2 = filters.smoother_loop(d, pp, kf2, full_state=False)
    print("vs library path: lml %.3e P %.3e var %.3e mu %.3e" % (abs(float(lml - lml2)) / abs(float(lml2)),
          rel(kf['P'], kf2['P'].cpu().numpy()), rel(var, var2.cpu().numpy()), rel(mu, mu2.cpu().numpy())))
