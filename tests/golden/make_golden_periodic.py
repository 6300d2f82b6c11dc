#!/usr/bin/env python
"""Golden vectors of the quasi-periodic / periodic state-space prior, produced by the reference's
own `ApproxSDEPeriodic_BN.{to_ss, expm}` (kernels/periodic.py:213-253) under the reference's own sequential filter /
smoother (`filter_loop`, `smoother_loop`), all executed in place from /root/reference on make_golden's numpy
stand-in for jax.  Two names of the periodic module come from libraries that are absent here and are supplied by
their scipy equivalents: `tfp.math.bessel_ive(v, z)` = `scipy.special.ive(v, z)` (exponentially scaled modified
Bessel function of the first kind) and `jax.scipy.linalg.expm` = `scipy.linalg.expm` (both Pade scaling-and-squaring).

    python tests/golden/make_golden_periodic.py      (needs /root/reference; writes tests/golden/periodic_*.npz)
"""
import importlib
import os
import sys
import types

import numpy as onp
import scipy.linalg as sla
import scipy.special as ssp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden_cvi import A  # noqa: E402

REF = mg.REF
Arr = mg.Arr

CASES = (  # name, n_terms, frequency, lengthscale, variance, extra Matern-3/2 (lengthscale, variance) summed on or None
    ("j3", 3, 2.0 * onp.pi / 1.3, 0.9, 1.4, None),
    ("j7", 7, 2.0 * onp.pi / 0.7, 0.6, 0.8, None),
    ("j10", 10, 2.0 * onp.pi / 2.1, 1.2, 1.1, None),
    ("j6_plus_m32", 6, 2.0 * onp.pi / 1.1, 0.8, 0.9, (0.7, 0.5)),
)


def main():
    assert os.path.isdir(REF)
    jax = mg.install_standin()
    sdes_mod = mg.install_package_tree()
    settings = sys.modules["stgp.settings"]
    settings.verbose = False
    settings.debug_mode = False
    jnp = jax.numpy
    kf = importlib.import_module("stgp.computation.filters.kalman_filter")
    rts = importlib.import_module("stgp.computation.filters.rts_smoother")
    ss = importlib.import_module("stgp.kernels.ss_utils")
    tfp = types.SimpleNamespace(math=types.SimpleNamespace(
        bessel_ive=lambda v, z: A(ssp.ive(onp.asarray(v, dtype=float), onp.asarray(z, dtype=float)))))
    ns = {"np": jnp, "jax": jax, "chex": sys.modules["chex"], "tfp": tfp, "expm": lambda M: A(sla.expm(onp.asarray(M)))}
    per = mg.extract("kernels/periodic.py", ["ApproxSDEPeriodic_BN.to_ss", "ApproxSDEPeriodic_BN.expm"], ns)
    # `extract` compiles each method in its own namespace, where the method's own name (`expm`) shadows the library
    # function of the same name it calls: point the global back at the library function
    per["ApproxSDEPeriodic_BN.expm"].__globals__["expm"] = ns["expm"]
    markov_Q = mg.extract("kernels/kernel.py", ["MarkovKernel.Q"], ns)["MarkovKernel.Q"]

    class PeriodicPrior(sdes_mod.LTI_SDE):
        """LTI_SDE facade over the reference's periodic kernel (+ optionally a Matern-3/2 summed on:
        kernel.py:134-160 -- block-diagonal F / Pinf / A, hstacked H)."""
        def __init__(self, order, freq, ls, var, extra):
            k = types.SimpleNamespace(order=order, n_terms=order, variance=var, lengthscale=ls, frequency=freq,
                                      include_dt=False, include_dt2=False, use_custom_bessel_ive=False)
            k.to_ss = lambda X_spatial=None: per["ApproxSDEPeriodic_BN.to_ss"](k, X_spatial)
            self.k, self.extra = k, extra

        def _parts(self):
            F, L, Qc, H, Pinf = self.k.to_ss()
            parts = [(onp.asarray(H), onp.asarray(Pinf))]
            if self.extra is not None:
                r = ss.matern32_temporal_state_space_rep(*self.extra)
                parts.append((onp.asarray(r[3]), onp.asarray(r[5])))
            return parts

        def m_inf(self, x, X_s, t):
            return A(onp.zeros([self.P_inf(None, None, None).shape[0], 1]))

        def P_inf(self, x, X_s, t):
            return A(sla.block_diag(*[p[1] for p in self._parts()]))

        def H(self, x, X_s, t):
            return A(onp.hstack([p[0] for p in self._parts()]))

        def expm(self, X_s, dt):
            blocks = [onp.asarray(per["ApproxSDEPeriodic_BN.expm"](self.k, dt))]
            if self.extra is not None:
                blocks.append(onp.asarray(ss.matern32_temporal_expm(dt, self.extra[0])))
            return A(sla.block_diag(*blocks))

        def Q(self, dt, A_k, P_inf, X_spatial=None):
            return markov_Q(None, dt, A_k, P_inf)                                   # kernel.py:207-209

    written = []
    for name, order, freq, ls, var, extra in CASES:
        for jitter in ((0.0, 1e-5) if name == "j3" else (1e-5,)):
            settings.jitter = jitter
            rng = onp.random.default_rng(1200 + order)
            T = 60
            t = onp.cumsum(rng.uniform(0.5, 1.5, T) * 0.05)
            Y = (onp.sin(freq * t) + 0.4 * onp.cos(2 * freq * t))[:, None] + 0.1 * rng.normal(size=(T, 1))
            Y[rng.uniform(size=Y.shape) < 0.15] = onp.nan
            R = onp.tile(0.02 * onp.eye(1), [T, 1, 1])
            prior = PeriodicPrior(order, freq, ls, var, extra)
            data = types.SimpleNamespace(X_time=A(t), X_space=None, Nt=T, Ns=1, P=1, Y_st=A(Y[:, :, None]))
            lml, res = kf.filter_loop(data, prior, R=A(R), filter_type="sequential")
            out = {"t": t, "Y": Y, "R": R, "jitter": jitter, "n_terms": order, "frequency": freq, "lengthscale": ls,
                   "variance": var, "extra_m32": onp.asarray(extra if extra is not None else [], dtype=float),
                   "H": onp.asarray(prior.H(None, None, None)), "P_inf": onp.asarray(prior.P_inf(None, None, None)),
                   "A_dt": onp.stack([onp.asarray(prior.expm(None, x)) for x in (0.0, 0.05, 0.9)]),
                   "lml": float(lml), "mf": onp.asarray(res["m"]), "Pf": onp.asarray(res["P"])}
            for fs in (False, True):
                mu, var_ = rts.smoother_loop(data, prior, res, full_state=fs, filter_type="sequential")
                out["ms_full%d" % fs], out["Ps_full%d" % fs] = onp.asarray(mu), onp.asarray(var_)
            fn = os.path.join(HERE, "periodic_%s_jit%s.npz" % (name, "1e-5" if jitter else "0"))
            onp.savez_compressed(fn, **out)
            written.append(fn)
    for f in written:
        print("wrote", os.path.relpath(f, HERE), os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
