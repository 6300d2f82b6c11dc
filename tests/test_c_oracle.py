"""The C oracle (oracle/ssm_oracle.c) against the numpy oracle (oracle/filters.py)."""
import numpy as np
import pytest

from oracle import c_oracle, filters, sde
from tests import synth


@pytest.mark.parametrize("s,nblk,fso", [(2, 1, False), (3, 1, False), (4, 1, False), (4, 2, False),
                                        (2, 3, False), (3, 2, True), (4, 1, True)])
def test_c_oracle_matches_numpy_oracle(s, nblk, fso):
    c_oracle.build()
    rng = np.random.default_rng(10 * s + nblk)
    B, T = 3, 150
    kind = {2: sde.Matern32, 3: sde.Matern52, 4: sde.Matern72}[s]
    ls = synth.log_uniform(rng, 0.5, 2.0, (B, nblk))
    var = synth.log_uniform(rng, 0.5, 2.0, (B, nblk))
    t = synth.time_grid(T, 0.1, rng)
    d = s * nblk
    priors = []
    for b in range(B):
        parts = [kind(ls[b, i], var[b, i]) for i in range(nblk)]
        if fso:
            priors.append(sde.LTI_SDE_Full_State_Obs(parts))
        else:
            priors.append(sde.LTI_SDE([parts[0] if nblk == 1 else sde.SumKernel(parts)]))
    H = priors[0].H()
    m = H.shape[0]
    Y = synth.noisy_series(B, T, m, rng, 0.1)
    R = synth.random_spd(rng, (B, T), m)
    lam = np.sqrt(2.0 * s - 1.0) / ls
    Pinf = np.stack([p.P_inf() for p in priors])
    out = c_oracle.filter_smooth(s, lam, Pinf, H, t, Y, R, jitter=1e-5, full_state=True, nthreads=2)
    for b in range(B):
        lml, mf, Pf, _ = filters.filter_sequential(priors[b], t, Y[b], R[b])
        ms, Ps = filters.smoother_sequential(priors[b], t, mf, Pf, full_state=True)
        assert abs(out["lml"][b] - lml) < 1e-11 * abs(lml)
        for a, r in ((out["mf"][b], mf[:, :, 0]), (out["Pf"][b], Pf), (out["ms"][b], ms[:, :, 0]), (out["Ps"][b], Ps)):
            assert np.abs(a - r).max() < 1e-11 * np.abs(r).max()
