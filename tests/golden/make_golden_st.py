#!/usr/bin/env python
"""Fourth golden generator: the reference's SEPARABLE spatio-temporal prior (BASELINE config 2), executed in place.

`SpatioTemporalSeperableKernel.to_ss / expm / Q` (kernels/kernel.py:213-265) and `space_time_state_space_rep`
(kernels/ss_utils.py:41-53) from /root/reference on the numpy stand-in of make_golden.py, driven through the reference's
own `filter_loop` / `smoother_loop` (sequential).  The temporal kernel is the reference's Matern closed form
(make_golden.Kern); the spatial kernel is an RBF Gram matrix evaluated here (spatial kernels are outside the hot path:
the product takes K_spatial as an input).  Vectors pin oracle/sde.py:SpaceTimeSeparable and the hand-written
separable-prior CUDA kernels (csrc/physs_kron.cu) to REFERENCE output, not only to the oracle.

    python tests/golden/make_golden_st.py      (needs /root/reference; writes tests/golden/st_*.npz)
"""
import importlib
import os
import sys
import types

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

REF = mg.REF
Arr = mg.Arr

# name: (temporal kernel (kind, ls, var), Ns, T, nan_frac, spatial lengthscale, seed)
ST_CASES = {
    "m32_ns18": (("m32", 0.8, 1.2), 18, 12, 0.1, 0.3, 11),
    "m52_ns12": (("m52", 0.6, 0.9), 12, 10, 0.15, 0.25, 12),
}


def main():
    assert os.path.isdir(REF)
    jax = mg.install_standin()
    sdes_mod = mg.install_package_tree()
    settings = sys.modules["stgp.settings"]
    settings.verbose = False
    jnp = jax.numpy
    kf = importlib.import_module("stgp.computation.filters.kalman_filter")
    rts = importlib.import_module("stgp.computation.filters.rts_smoother")
    ss = importlib.import_module("stgp.kernels.ss_utils")
    Kern, _ = mg.build_prior_classes(jax, sdes_mod)
    ns = {"np": jnp, "jax": jax, "chex": sys.modules["chex"], "space_time_state_space_rep": ss.space_time_state_space_rep}
    ref = mg.extract("kernels/kernel.py", ["SpatioTemporalSeperableKernel.to_ss", "SpatioTemporalSeperableKernel.expm",
                                           "SpatioTemporalSeperableKernel.Q"], ns)

    class RBF:
        def __init__(self, ls):
            self.ls = ls

        def K(self, X1, X2):                       # called with the dummy time column the reference prepends
            a, b = onp.asarray(X1)[:, 1:], onp.asarray(X2)[:, 1:]
            d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
            return (onp.exp(-0.5 * d2 / self.ls ** 2) + 1e-6 * onp.eye(a.shape[0]) * (a.shape == b.shape)).view(Arr)

    class STKernel:
        """Carrier of the attributes the reference methods read (kernel.py:218-224)."""
        def __init__(self, k1, k2):
            self.k1, self.k2 = k1, k2
            self.spatial_output_dim, self.whiten_space, self.stationary = 1, False, True

    class Prior(sdes_mod.LTI_SDE):
        """LTI_SDE over one separable latent: every quantity from the reference methods above (sdes.py:58-97)."""
        def __init__(self, kern):
            self.kern = kern

        def _ss(self, X_s):
            return ref["SpatioTemporalSeperableKernel.to_ss"](self.kern, X_s)

        def m_inf(self, x, X_s, t):
            return onp.asarray(self._ss(X_s)[4]).view(Arr)

        def P_inf(self, x, X_s, t):
            return onp.asarray(self._ss(X_s)[5]).view(Arr)

        def H(self, x, X_s, t):
            return onp.asarray(self._ss(X_s)[3]).view(Arr)

        def expm(self, X_s, dt):
            return onp.asarray(ref["SpatioTemporalSeperableKernel.expm"](self.kern, dt, X_s)).view(Arr)

        def Q(self, dt, A_k, P_inf, X_spatial=None):
            return ref["SpatioTemporalSeperableKernel.Q"](self.kern, dt, A_k, P_inf, X_spatial)

    written = []
    for name, (tk, Ns, T, nan_frac, ls_s, seed) in ST_CASES.items():
        for jitter in (1e-5, 0.0):
            settings.jitter = jitter
            rng = onp.random.default_rng(seed)
            Xs = rng.uniform(size=(Ns, 2))
            prior = Prior(STKernel(Kern(*tk), RBF(ls_s)))
            t, Y, R = mg.synth(T, Ns, nan_frac, rng)
            data = types.SimpleNamespace(X_time=t.view(Arr), X_space=Xs.view(Arr), Nt=T, Ns=Ns, P=1,
                                         Y_st=Y[:, None, :].view(Arr))
            Ks = onp.asarray(prior.kern.k2.K(onp.hstack([onp.zeros((Ns, 1)), Xs]), onp.hstack([onp.zeros((Ns, 1)), Xs])))
            out = {"t": t, "Y": Y, "R": R, "jitter": jitter, "Xs": Xs, "Ks": Ks, "temporal": onp.array(tk[1:]),
                   "P_inf": onp.asarray(prior.P_inf(None, Xs.view(Arr), None)), "H": onp.asarray(prior.H(None, Xs.view(Arr), None)),
                   "A_dt": onp.stack([onp.asarray(prior.expm(Xs.view(Arr), x)) for x in (0.0, 0.05, 0.9)])}
            lml, res = kf.filter_loop(data, prior, R=R.view(Arr), filter_type="sequential")
            out["seq_lml"], out["seq_mf"], out["seq_Pf"] = float(lml), onp.asarray(res["m"]), onp.asarray(res["P"])
            for fs in (False, True):
                mu, var = rts.smoother_loop(data, prior, res, full_state=fs, filter_type="sequential")
                out["seq_ms_full%d" % fs], out["seq_Ps_full%d" % fs] = onp.asarray(mu), onp.asarray(var)
            fn = os.path.join(HERE, "st_%s_jit%s.npz" % (name, "1e-5" if jitter else "0"))
            onp.savez_compressed(fn, **out)
            written.append(fn)
    for fn in written:
        print("wrote", os.path.relpath(fn, HERE), os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()
