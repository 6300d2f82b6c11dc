import sys, numpy as np, torch
sys.path.insert(0, '.')
from tests.test_gpu_pscan import _batch_problem, rel
from physs_gp_b200 import ops
dev = torch.device('cuda:0')
B,T,d,m,given,tm,chunk,warm = 40,2500,2,1,False,True,97,97
dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s = _batch_problem(dev, B, T, d, m, given, 17 + d, tm)
lml, mf, Pf = ops.kf_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter=1e-5)
ms, Ps = ops.rts_smooth(dt_s, mf, Pf, disc_s, jitter=1e-5)
for polish in (1,3,6,10):
    lml2, mf2, Pf2, st = ops.pscan_filter_spec(dt_f, Y, R, H, m0, P0, disc_f, chunk_len=chunk, warm=warm, jitter=1e-5, polish=polish)
    ms2, Ps2, st2 = ops.pscan_smooth_spec(dt_s, mf, Pf, disc_s, chunk_len=chunk, warm=warm, jitter=1e-5, polish=polish)
    print(polish, int(st.item()), int(st2.item()), rel(lml2, lml.cpu().numpy()), rel(mf2, mf.cpu().numpy()), rel(Pf2, Pf.cpu().numpy()), rel(ms2, ms.cpu().numpy()), rel(Ps2, Ps.cpu().numpy()))
    e=((ms2-ms).abs().amax(dim=(0,2))/ms.abs().max()).cpu().numpy()
    print('  worst smoother step', int(e.argmax()), e.max(), 'chunk pos', int(e.argmax())%chunk)
