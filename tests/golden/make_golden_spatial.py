#!/usr/bin/env python
"""Golden vectors of the spatial conditional: the reference's own
`gaussian_spatial_conditional_cholesky` (computation/marginals.py:82-113), vmapped over time steps on factors formed by
the reference's own `cholesky(add_jitter(.))` (computation/matrix_ops.py:108-110, 234-236) exactly as the f_only branch
of `spatial_conditional_block` does (computation/spatial_conditionals.py:137-207), executed in place from
/root/reference on make_golden's numpy stand-in for jax.  The Gram matrices (inputs) are squared-exponential /
Matern-3/2 evaluations written here -- spatial kernels are outside the path.

    python tests/golden/make_golden_spatial.py      (needs /root/reference; writes tests/golden/spatial_cond_*.npz)
"""
import importlib
import os
import sys

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden_cvi import A  # noqa: E402

REF = mg.REF

CASES = (  # name, M (inducing points), N (new points), T, spatial dim, jitter, time-varying Ktt
    ("m12_n7", 12, 7, 9, 1, 1e-5, False),
    ("m40_n33", 40, 33, 6, 2, 1e-5, True),
    ("m75_n50", 75, 50, 4, 2, 1e-6, False),
)


def gram(X1, X2, ls, var, kind):
    r = onp.sqrt(((X1[:, None, :] - X2[None, :, :]) ** 2).sum(-1))
    if kind == "rbf":
        return var * onp.exp(-0.5 * (r / ls) ** 2)
    a = onp.sqrt(3.0) * r / ls
    return var * (1.0 + a) * onp.exp(-a)


def main():
    assert os.path.isdir(REF)
    jax = mg.install_standin()
    mg.install_package_tree()
    settings = sys.modules["stgp.settings"]
    jnp = jax.numpy
    mops = importlib.import_module("stgp.computation.matrix_ops")
    ns = {"np": jnp, "jax": jax, "chex": sys.modules["chex"], "jit": mg._jit, "triangular_solve": mops.triangular_solve,
          "settings": settings, "cholesky": mops.cholesky, "add_jitter": mops.add_jitter}
    fn = mg.extract("computation/marginals.py", ["gaussian_spatial_conditional_cholesky"], ns)[
        "gaussian_spatial_conditional_cholesky"]
    written = []
    for name, M, N, T, D, jitter, varying in CASES:
        rng = onp.random.default_rng(4000 + M)
        X = rng.uniform(0.0, 1.0, [M, D])
        XS = rng.uniform(-0.1, 1.1, [N, D])
        kind = "rbf" if M == 12 else "m32"
        ls, var = 0.35, 1.3
        Kzz, Ksz, Kss = gram(X, X, ls, var, kind), gram(XS, X, ls, var, kind), gram(XS, XS, ls, var, kind)
        Ktt = (0.7 + 0.3 * rng.uniform(size=T)) if varying else onp.full(T, 0.9)
        # a per-step posterior at the inducing points: SPD, smaller than the prior
        pm = rng.normal(size=[T, M, 1])
        pv = onp.zeros([T, M, M])
        for t in range(T):
            B = rng.normal(size=[M, M]) * 0.2
            pv[t] = 0.3 * Kzz + B @ B.T / M + 0.01 * onp.eye(M)
        settings.jitter = jitter
        # spatial_conditionals.py:137-141, 146: factors through the reference's cholesky(add_jitter(.))
        S_chol = jax.vmap(lambda S: mops.cholesky(mops.add_jitter(S, settings.jitter)), 0)(A(pv))
        Kzz_chol = mops.cholesky(mops.add_jitter(A(Kzz), settings.jitter))
        mean_x, mean_xs = A(onp.zeros([M, 1])), A(onp.zeros([N, 1]))
        # spatial_conditionals.py:181-194: vmap over (Ktt_full, pred_mean, S)
        Ktt_full = onp.stack([k * onp.ones([N, N]) for k in Ktt])
        mu, sig = jax.vmap(fn, [None, None, None, None, None, 0, 0, 0, None, None])(
            A(XS), A(X), Kzz_chol, A(Ksz), A(Kss), A(Ktt_full), A(pm), S_chol, mean_x, mean_xs)
        out = {"X": X, "XS": XS, "Kzz": Kzz, "Ksz": Ksz, "Kss": Kss, "Ktt": Ktt, "pred_mean": pm, "pred_var": pv,
               "jitter": jitter, "mu": onp.asarray(mu), "var": onp.asarray(sig)[:, None]}
        f = os.path.join(HERE, "spatial_cond_%s.npz" % name)
        onp.savez_compressed(f, **out)
        written.append(f)
    for f in written:
        print("wrote", os.path.relpath(f, HERE), os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
