"""CPU tests of the host-side logic that needs no GPU: chunk-length choice, the prior's hyper-parameter chain
rule, the likelihood containers on the merged train + test grid of predict_f."""
import numpy as np
import pytest
import torch

from physs_gp_b200 import likelihood, sdes


def _even_chunk_len():
    # ops imports the ctypes binding lazily; the helper itself is pure Python
    from physs_gp_b200 import ops
    return ops.even_chunk_len


def test_even_chunk_len_divides_or_falls_back():
    f = _even_chunk_len()
    assert f(1000000, 256) == 250 and 1000000 % f(1000000, 256) == 0
    assert f(500000, 256) == 250
    assert f(10007, 256) == 256                      # prime length: nothing within +-25 % divides it
    assert f(100, 256) == 256 and f(512, 256) == 256
    for T, L in ((123456, 300), (99999, 128), (65536, 200)):
        c = f(T, L)
        assert c == L or (T % c == 0 and abs(c - L) <= 0.25 * L + 1)


@pytest.mark.parametrize("s", [1, 2, 3, 4])
def test_hyper_grads_is_the_chain_rule_of_the_closed_forms(s):
    """<glam, d lam> + <gPinf, d Pinf> for perturbations of (lengthscale, variance), by central differences."""
    rng = np.random.default_rng(s)
    B, nblk = 3, 2
    ls = rng.uniform(0.5, 1.5, (B, nblk))
    var = rng.uniform(0.5, 1.5, (B, nblk))
    prior = sdes.BatchedMaternSDE(s, ls, var)
    d = s * nblk
    glam = rng.normal(size=(B, nblk))
    gP = rng.normal(size=(B, d, d))
    g_ls, g_var = prior.hyper_grads(glam, gP)

    def obj(ls_, var_):
        p = sdes.BatchedMaternSDE(s, ls_, var_)
        return np.sum(glam * p.lam(), axis=1) + np.sum(gP * p.P_inf(), axis=(1, 2))      # per series

    h = 1e-6
    for b in range(nblk):
        e = np.zeros((B, nblk)); e[:, b] = h
        fd_ls = (obj(ls + e, var) - obj(ls - e, var)) / (2 * h)
        fd_var = (obj(ls, var + e) - obj(ls, var - e)) / (2 * h)
        np.testing.assert_allclose(g_ls[:, b], fd_ls, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(g_var[:, b], fd_var, rtol=1e-6, atol=1e-8)


def test_R_predict_on_the_merged_grid():
    Nt, NS, m = 5, 3, 2
    t = np.array([0.0, 1.0, 2.0, 3.0, 4.0])
    ts = np.array([2.0, 0.5, 9.0])                    # one duplicate of a training time
    _, ui, ri = np.unique(np.concatenate([t, ts]), return_index=True, return_inverse=True)
    g = likelihood.Gaussian(0.3)
    assert g.R_predict(Nt, NS, ui, m).shape == (1, m, m)
    V = np.arange(Nt)[:, None, None] + np.eye(m)[None] * 10.0
    for Vv in (V, torch.as_tensor(V)):
        R = likelihood.BlockDiagonalGaussian(Vv).R_predict(Nt, NS, ui, m)
        R = R.numpy() if isinstance(R, torch.Tensor) else R
        assert R.shape == (len(ui), m, m)
        # training rows keep their own block (the duplicate test time 2.0 keeps the TRAINING block), test rows get I
        merged_t = np.concatenate([t, ts])[ui]
        for k, tk in enumerate(merged_t):
            if tk in t:
                np.testing.assert_array_equal(R[k], V[list(t).index(tk)])
            else:
                np.testing.assert_array_equal(R[k], np.eye(m))
    # the test rows are found again through the inverse index
    assert list(np.concatenate([t, ts])[ui][ri.reshape(-1)[Nt:]]) == list(ts)


def test_get_R_R_inv_follows_the_likelihood_class():
    """sde_gp.py:30-43: a PrecisionBlockDiagonalGaussian hands (None, precision) to filter_loop, every other Gaussian
    container (variance, None)."""
    import numpy as np
    from physs_gp_b200 import likelihood
    P = np.tile(2.0 * np.eye(3), [5, 1, 1])
    R, R_inv = likelihood.get_R_R_inv(likelihood.PrecisionBlockDiagonalGaussian(P), 5, 3)
    assert R is None and R_inv is P
    R, R_inv = likelihood.get_R_R_inv(likelihood.BlockDiagonalGaussian(P), 5, 3)
    assert R is P and R_inv is None
    R, R_inv = likelihood.get_R_R_inv(likelihood.Gaussian(0.3), 5, 2)
    assert R_inv is None and R.shape == (1, 2, 2) and R[0, 1, 1] == 0.3


def test_packed_route_declines_without_touching_the_gpu():
    """filters.filter_smooth_fused returns None -- the caller then takes filter_loop + smoother_loop -- for everything the
    packed call does not cover, and decides that before any device work: precision sites, collocation priors, state
    dims above 4, the switch settings.fused_packed."""
    import numpy as np
    from physs_gp_b200 import data, filters, sdes, settings
    t = np.linspace(0.0, 1.0, 20)
    Y = np.zeros((40, 20, 1, 1))
    d = data.TemporalData(t, Y)
    big = sdes.BatchedMaternSDE(4, np.ones((40, 2)))                     # state dim 8
    assert filters.filter_smooth_fused(d, big, R=np.eye(1)[None]) is None
    small = sdes.BatchedMaternSDE(4, np.ones((40, 1)))
    assert filters.filter_smooth_fused(d, small, R=None, R_inv=np.eye(1)[None]) is None
    old = settings.fused_packed
    settings.fused_packed = False
    try:
        assert filters.filter_smooth_fused(d, small, R=np.eye(1)[None]) is None
    finally:
        settings.fused_packed = old
