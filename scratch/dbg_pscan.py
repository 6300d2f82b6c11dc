import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_gpu_pscan import _batch_problem, rel
from physs_gp_b200 import ops
dev = torch.device('cuda:0')
B,T,d,m,given,tm,chunk = 1,30000,12,12,False,False,64
dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(dev, B, T, d, m, given, 7 + d, tm)
for jitter in (1e-5, 0.0):
    lml, mf, Pf = ops.kf_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter=jitter)
    print("jitter", jitter, "P scale", float(Pf.abs().max()), "P0 scale", float(P0.abs().max()), "m scale", float(mf.abs().max()))
    for polish in (0,1,2,3):
        for delta in (1e-10, 1e-8):
            lml2, mf2, Pf2, st = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=chunk, jitter=jitter, polish=polish, delta=delta, return_status=True)
            em = (mf2-mf).abs().amax(dim=(0,2)); eP=(Pf2-Pf).abs().amax(dim=(0,2,3))
            print(" polish",polish,"delta",delta,"status",int(st.item()),"rel m",rel(mf2,mf.cpu().numpy()),"rel P",rel(Pf2,Pf.cpu().numpy()), "worst step m", int(em.argmax())%chunk, "worst step P", int(eP.argmax())%chunk)
    # per-step profile of P error in a chunk after polish=0
    lml2, mf2, Pf2 = ops.pscan_filter(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=chunk, jitter=jitter, polish=0)
    eP=((Pf2-Pf).abs().amax(dim=(0,2,3))/Pf.abs().amax()).cpu().numpy()
    print(" profile chunk 5:", ["%.1e"%x for x in eP[5*chunk:5*chunk+12]])
