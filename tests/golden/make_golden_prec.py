#!/usr/bin/env python
"""Golden generator for the PRECISION parameterisation of the CVI sites ('NG_Precision'): the reference's own
`theta_precision_to_lambda` / `lambda_to_theta_precision` (computation/natural_gradients/
exponential_family_transforms.py:44-53,85-95, imported as the real module) around its own `cvi_block_update`
(cvi_nat_grad.py:47-87, ast-extracted) -- the composition `natural_gradients(VGP, FullConjugateGaussian,
"NG_Precision")` performs (cvi_parameterisations.py:95-113 with cvi_nat_grad_utils.py:62-63) -- and `mat_inv`
(computation/matrix_ops.py:383-385), which turns the stored precision into the surrogate likelihood's variance
(likelihood/gaussian.py:96-105).  Everything runs in place from /root/reference on the numpy stand-in of
make_golden.py; nothing is copied.

    python tests/golden/make_golden_prec.py       (needs /root/reference; writes tests/golden/cvi_blocks_prec.npz)
"""
import importlib
import os
import sys

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

Arr = mg.Arr


def main():
    jax = mg.install_standin()
    mg.install_package_tree()
    settings = importlib.import_module("stgp.settings")
    eft = importlib.import_module("stgp.computation.natural_gradients.exponential_family_transforms")
    mo = importlib.import_module("stgp.computation.matrix_ops")
    ns = {"np": jax.numpy, "chex": sys.modules["chex"], "jit": mg._jit, "jax": jax, "settings": settings,
          "partial": __import__("functools").partial}
    for k in ("cholesky", "cholesky_solve", "add_jitter"):
        ns[k] = getattr(mo, k)
    blk = mg.extract("computation/natural_gradients/cvi_nat_grad.py", ["cvi_block_update"], ns)["cvi_block_update"]
    rng = onp.random.default_rng(23)
    out = {}
    for D in (1, 2, 3, 4, 6, 8):
        for ngj in (1e-7, 1e-5):
            settings.ng_jitter = ngj
            settings.jitter = 1e-5
            G = rng.normal(size=(D, D))
            Lam = G @ G.T + 0.5 * onp.eye(D)                 # the stored site precision
            Yt = rng.normal(size=(D, 1))
            l1, l2 = eft.theta_precision_to_lambda(Yt.view(Arr), Lam.view(Arr))
            G2 = rng.normal(size=(D, D))
            S = G2 @ G2.T + 0.2 * onp.eye(D)
            mq = rng.normal(size=(D, 1))
            dm = rng.normal(size=(D, 1))
            G3 = rng.normal(size=(D, D))
            dS = -(G3 @ G3.T) * 0.3
            beta = 0.37
            n1, n2 = blk(l1, l2, mq.view(Arr), S.view(Arr), dm.view(Arr), dS.view(Arr), beta, None)
            t1, t2 = eft.lambda_to_theta_precision(n1, n2)
            var = mo.mat_inv(Lam.view(Arr))                   # PrecisionBlockDiagonalGaussian.variance
            key = "D%d_ngj%s" % (D, "1e-7" if ngj == 1e-7 else "1e-5")
            for nm, val in (("Lam", Lam), ("Yt", Yt), ("l1", l1), ("l2", l2), ("S", S), ("mq", mq), ("dm", dm),
                            ("dS", dS), ("beta", beta), ("n1", n1), ("n2", n2), ("t1", t1), ("t2", t2),
                            ("var", var), ("ng_jitter", ngj)):
                out["%s_%s" % (key, nm)] = onp.asarray(val)
    fn = os.path.join(HERE, "cvi_blocks_prec.npz")
    onp.savez_compressed(fn, **out)
    print("wrote", os.path.relpath(fn, HERE), os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()
