import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_fused import _packed_problem
from physs_gp_b200 import ops
dev = torch.device('cuda:0')
for shape in [(4,1),(2,2),(1,4),(2,1),(1,2)]:
    s, nblk = shape
    args = _packed_problem(dev, np.random.default_rng(11), 70, 121, s, nblk, "matern")
    dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s = args
    a = ops.kf_filter_smooth(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H, jitter=1e-5)
    a2 = ops.kf_filter_smooth(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H, jitter=1e-5)
    b = ops.kf_filter_smooth_packed(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H, jitter=1e-5)
    b2 = ops.kf_filter_smooth_packed(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=H, jitter=1e-5)
    f = ops.kf_filter_smooth(dt_f, dt_s, Yt, R, H, m0, P0, disc_f, disc_s, Hout=None, jitter=1e-5)
    Hm = H.reshape(-1, H.shape[-1])
    pm = torch.einsum('ad,btd->bta', Hm, f[3]); pP = torch.einsum('ad,btde,ce->btac', Hm, f[4], Hm)
    Pf = f[2]
    print(shape, 'two-call repeat equal', torch.equal(a[3], a2[3]), 'packed repeat equal', torch.equal(b[1], b2[1]),
          'max|two-packed| m', float((a[3]-b[1]).abs().max()), 'P', float((a[4]-b[2]).abs().max()),
          'two vs torch proj', float((a[3]-pm).abs().max()), 'packed vs torch proj', float((b[1]-pm).abs().max()),
          'Pf asym', float((Pf - Pf.transpose(-1,-2)).abs().max()))
    nz = (a[3]-b[1]).abs().amax(dim=(0,2)).nonzero().flatten()
    print('  steps with diffs', nz[:10].tolist(), '... count', nz.numel())
