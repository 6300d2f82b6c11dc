#!/usr/bin/env python
"""Third golden generator: the reference's collocation (EKF) filter step, executed in place.

Runs `filter_loop` -> `evoke('filter', 'sequential')` -> `evoke('kf_predict_step', <PDE>, 'sequential')`
(computation/filters/kalman_filter.py:487-547, 439-485, 340-427) and `smoother_loop` -> `rts_step_wrapper(PDE)`
(rts_smoother.py:108-150) from /root/reference on the numpy stand-in of make_golden.py.  The model is a bare
subclass of the stand-in `PDE` type (so the reference's dispatch picks the PDE step) carrying what that step
reads: `.parent` (the LTI prior assembled from the reference's Matern closed forms), `.H`, `.forward_g`,
`.H_jac` = the reference's own `PDE.jac` (transforms/pdes.py:236-245) with `jax.jacfwd` supplied by complex-step
differentiation, `.boundary_conditions`, `.psuedo_observations`, `.observe_data`.  Residuals: the reference's
`DampedPendulum1D.forward` (pdes.py:584-597) and `Pendulum1D.forward`, plus a cubic reaction term with a
time-dependent forcing written here (the step code under test is the reference's either way).

    python tests/golden/make_golden_ekf.py      (needs /root/reference; writes tests/golden/ekf_*.npz)
"""
import importlib
import os
import sys
import types

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden_cvi import A, jacfwd_complex  # noqa: E402

REF = mg.REF
Arr = mg.Arr


def main():
    assert os.path.isdir(REF)
    jax = mg.install_standin()
    sdes_mod = mg.install_package_tree()
    settings = sys.modules["stgp.settings"]
    settings.verbose = False
    settings.debug_mode = False
    jnp = jax.numpy
    jax.jacfwd = lambda f, argnums=0: jacfwd_complex(f, argnums)
    pdes_mod = sys.modules["stgp.transforms.pdes"]
    kf = importlib.import_module("stgp.computation.filters.kalman_filter")
    rts = importlib.import_module("stgp.computation.filters.rts_smoother")
    Kern, Prior = mg.build_prior_classes(jax, sdes_mod)
    chex = sys.modules["chex"]
    ns = {"np": jnp, "jax": jax, "chex": chex}
    pde_jac = mg.extract("transforms/pdes.py", ["PDE.jac"], ns)["PDE.jac"]
    damped = mg.extract("transforms/pdes.py", ["DampedPendulum1D.forward"], ns)["DampedPendulum1D.forward"]
    simple = mg.extract("transforms/pdes.py", ["Pendulum1D.forward"], ns)["Pendulum1D.forward"]

    class LTIParent(Prior):
        """LTI_SDE facade: adds state_space_representation (sdes.py:58-60) to make_golden's prior"""
        def state_space_representation(self, X_s, dt, t):
            return (None, None, None, self.H(None, X_s, None), self.m_inf(None, X_s, None),
                    self.P_inf(None, X_s, None))

        def state_space_dim(self):
            return self.P_inf(None, None, None).shape[0]
        spatial_output_dim = 1

    class Model(pdes_mod.PDE):
        def __init__(self, parent, g_fn, boundary, observe_data, y_pseudo):
            self.parent, self._g = parent, g_fn
            self.boundary_conditions, self.observe_data, self._yp = boundary, observe_data, y_pseudo

        def m_inf(self, x, X_s, t):
            return self.parent.m_inf(x, X_s, t)

        def P_inf(self, x, X_s, t):
            return self.parent.P_inf(x, X_s, t)

        def H(self, x, X_s, t):
            return self.parent.H(None, X_s, None)

        def H_full_state(self, x, X_s, t):
            return onp.eye(self.parent.state_space_dim()).view(Arr)

        def forward_g(self, x, X_s, t):
            return self._g(x, t)

        def jac(self, x, X_s, t):
            return pde_jac(self, x, X_s, t)

        def H_jac(self, x, X_s, t):                       # pdes.py:244-245
            return self.jac(x, X_s, t)

        def psuedo_observations(self, X_s):
            return onp.array(self._yp, dtype=float)[:, None].view(Arr)

    pend = types.SimpleNamespace(g_param=types.SimpleNamespace(value=9.81), l_param=types.SimpleNamespace(value=1.3),
                                 b_param=types.SimpleNamespace(value=0.35))
    a_, b_ = 9.81 / 1.3, 0.35
    nan = onp.nan
    # The reference's step needs >= 2 collocation outputs (`np.squeeze(f)[..., None]`, kalman_filter.py:413, is
    # rank 1 for a single one and trips chex.assert_rank in log_gaussian_with_mask): every case has two, one of
    # which may be switched off by a NaN pseudo-observation, as SimpleODE does (pdes.py:477-478).
    # name: (kernel, residual fn of (x [d,1], t) -> [2], descriptors for the tests, boundary?, observe_data, y_pseudo)
    cases = {
        "damped_m72": (("m72", 0.6, 2.0),
                       lambda x, t: jnp.hstack([damped(pend, x[:3, 0]), x[3, 0] + b_ * x[2, 0] + a_ * x[1, 0]]),
                       [dict(w=[0.0, b_, 1.0, 0.0], terms=[("sin", 0, a_)]), dict(w=[0.0, a_, b_, 1.0], terms=[])],
                       False, True, [0.0, 0.0]),
        "pendulum_m52_boundary": (("m52", 0.8, 1.5),
                                  lambda x, t: jnp.hstack([simple(pend, x[:3, 0]), x[1, 0]]),
                                  [dict(w=[0.0, 0.0, 1.0], terms=[("sin", 0, a_)]), dict(w=[0.0, 1.0, 0.0], terms=[])],
                                  True, True, [0.0, nan]),
        "cubic_forced_m32": (("m32", 0.9, 1.2),
                             lambda x, t: jnp.array([x[1, 0] + 0.7 * x[0, 0] ** 3 - x[0, 0] - jnp.sin(t),
                                                     x[0, 0] ** 2 + jnp.cos(x[1, 0])]),
                             [dict(w=[-1.0, 1.0], terms=[("cube", 0, 0.7)], forcing="-sin(t)"),
                              dict(w=[0.0, 0.0], terms=[("square", 0, 1.0), ("cos", 1, 1.0)])],
                             False, False, [0.0, nan]),
        "damped_m72_no_colloc": (("m72", 0.6, 2.0),
                                 lambda x, t: jnp.hstack([damped(pend, x[:3, 0]), x[3, 0]]),
                                 [dict(w=[0.0, b_, 1.0, 0.0], terms=[("sin", 0, a_)]), dict(w=[0.0, 0.0, 0.0, 1.0], terms=[])],
                                 False, True, [nan, nan]),
    }
    written = []
    for jitter in (1e-5, 0.0):
        settings.jitter = jitter
        for name, (kern, g_fn, desc, has_bnd, observe, yp) in cases.items():
            rng = onp.random.default_rng(sum(map(ord, name)))
            parent = LTIParent([[Kern(*kern)]], False)
            T = 40
            t = onp.cumsum(rng.uniform(0.5, 1.5, T) * 0.05)
            Y = 0.8 * onp.cos(2.5 * t)[:, None] + 0.05 * rng.normal(size=(T, 1))
            Y[rng.uniform(size=Y.shape) < 0.3] = onp.nan
            R = onp.tile(0.05 ** 2 * onp.eye(1), [T, 1, 1])
            bnd = None
            if has_bnd:
                bnd = onp.full((T, 1, 1), onp.nan)
                bnd[0, 0, 0] = 0.8
            model = Model(parent, g_fn, None if bnd is None else A(bnd), observe, yp)
            data = types.SimpleNamespace(X_time=A(t), X_space=None, Nt=T, Ns=1, P=1, Y_st=A(Y[:, :, None]))
            try:
                lml, res = kf.filter_loop(data, model, R=A(R), filter_type="sequential")
            except onp.linalg.LinAlgError:
                # zero-noise updates with jitter = 0 (a masked pseudo-observation, or a state the boundary update has
                # already pinned) hand cholesky a singular matrix: the reference then returns NaN; nothing to pin
                print("  %s, jitter %g: singular innovation covariance in the reference (NaN) -- skipped" % (name, jitter))
                continue
            mu, var = rts.smoother_loop(data, model, res, full_state=True, filter_type="sequential")
            out = {"t": t, "Y": Y, "R": R, "jitter": jitter, "observe_data": observe, "y_pseudo": yp,
                   "kernel": onp.array([kern[0]]), "hyper": onp.array(kern[1:]), "n_res": len(desc),
                   "lml": float(lml), "mf": onp.asarray(res["m"]), "Pf": onp.asarray(res["P"]),
                   "ms": onp.asarray(mu), "Ps": onp.asarray(var)}
            for p, dsc in enumerate(desc):
                out["w%d" % p] = onp.array(dsc["w"])
                out["term_kind%d" % p] = onp.array([k for k, _, _ in dsc["terms"]], dtype="U8")
                out["term_idx%d" % p] = onp.array([i for _, i, _ in dsc["terms"]], dtype=int)
                out["term_coef%d" % p] = onp.array([c for _, _, c in dsc["terms"]], dtype=float)
                out["forcing%d" % p] = -onp.sin(t) if "forcing" in dsc else onp.zeros(0)
            if bnd is not None:
                out["boundary"] = bnd[:, :, 0]
            fn = os.path.join(HERE, "ekf_%s_jit%s.npz" % (name, "1e-5" if jitter else "0"))
            onp.savez_compressed(fn, **out)
            written.append(fn)
    # =============================================== systems of ODEs over several latents: the reference's own
    # LotkaVolterra components (transforms/pdes.py:912-1008, stacked as LotkaVolterra.forward does, :1059-1072) and the
    # three Lorenz components (:818-910) over 2 / 3 independent Matern-3/2 latents (state x, xt, y, yt[, z, zt]), one
    # observed output per latent.  Bilinear terms are stored as term_idx = i | (j << 8), kind "prod".
    lv_names = ["_LotkaVolterraSystemX.forward", "_LotkaVolterraSystemX._dfdt", "_LotkaVolterraSystemY.forward",
                "_LotkaVolterraSystemY._dfdt"]
    lv = mg.extract("transforms/pdes.py", lv_names, ns)
    lz = mg.extract("transforms/pdes.py", ["_LorenzSystemX.forward", "_LorenzSystemY.forward", "_LorenzSystemZ.forward"], ns)
    al, be, de, ga = 1.1, 0.4, 0.1, 0.4
    sg, rho, bt = 10.0, 28.0, 8.0 / 3.0
    par = lambda v: types.SimpleNamespace(value=v)
    lvx = types.SimpleNamespace(alpha_param=par(al), beta_param=par(be))
    lvx._dfdt = lambda f: lv["_LotkaVolterraSystemX._dfdt"](lvx, f)
    lvy = types.SimpleNamespace(delta_param=par(de), gamma_param=par(ga))
    lvy._dfdt = lambda f: lv["_LotkaVolterraSystemY._dfdt"](lvy, f)
    lzs = types.SimpleNamespace(sigma_param=par(sg), rho_param=par(rho), beta_param=par(bt))
    enc = lambda i, j: i | (j << 8)
    sys_cases = {
        "lotka_volterra": ([("m32", 1.5, 4.0), ("m32", 1.2, 3.0)],
                           lambda x, t: jnp.hstack([onp.squeeze(lv["_LotkaVolterraSystemX.forward"](lvx, x)),
                                                    onp.squeeze(lv["_LotkaVolterraSystemY.forward"](lvy, x))]),
                           [dict(w=[-al, 1.0, 0.0, 0.0], terms=[("prod", enc(0, 2), be)]),
                            dict(w=[0.0, 0.0, ga, 1.0], terms=[("prod", enc(0, 2), -de)])],
                           True, True, [0.0, 0.0]),
        "lorenz": ([("m32", 0.3, 60.0), ("m32", 0.3, 80.0), ("m32", 0.3, 90.0)],
                   lambda x, t: jnp.hstack([onp.squeeze(lz["_LorenzSystemX.forward"](lzs, x)),
                                            onp.squeeze(lz["_LorenzSystemY.forward"](lzs, x)),
                                            onp.squeeze(lz["_LorenzSystemZ.forward"](lzs, x))]),
                   [dict(w=[sg, 1.0, -sg, 0.0, 0.0, 0.0], terms=[]),
                    dict(w=[-rho, 0.0, 1.0, 1.0, 0.0, 0.0], terms=[("prod", enc(0, 4), 1.0)]),
                    dict(w=[0.0, 0.0, 0.0, 0.0, bt, 1.0], terms=[("prod", enc(0, 2), -1.0)])],
                   False, True, [0.0, 0.0, 0.0]),
    }
    settings.jitter = 1e-5
    for name, (kerns, g_fn, desc, has_bnd, observe, yp) in sys_cases.items():
        rng = onp.random.default_rng(sum(map(ord, name)))
        nl = len(kerns)
        parent = LTIParent([[Kern(*k)] for k in kerns], False)
        T = 40
        t = onp.cumsum(rng.uniform(0.5, 1.5, T) * (0.05 if nl == 2 else 0.004))
        base = onp.stack([2.0 + onp.sin(1.3 * t + q) for q in range(nl)], 1) if nl == 2 else \
            onp.stack([1.0 + 3.0 * t, 1.5 + 5.0 * t, 20.0 + 2.0 * t], 1)
        Y = base + 0.05 * rng.normal(size=(T, nl))
        Y[rng.uniform(size=Y.shape) < 0.3] = onp.nan
        R = onp.tile(0.05 ** 2 * onp.eye(nl), [T, 1, 1])
        bnd = None
        if has_bnd:
            bnd = onp.full((T, nl, 1), onp.nan)
            bnd[0, :, 0] = base[0]
        model = Model(parent, g_fn, None if bnd is None else A(bnd), observe, yp)
        data = types.SimpleNamespace(X_time=A(t), X_space=None, Nt=T, Ns=1, P=nl, Y_st=A(Y[:, :, None]))
        lml, res = kf.filter_loop(data, model, R=A(R), filter_type="sequential")
        mu, var = rts.smoother_loop(data, model, res, full_state=True, filter_type="sequential")
        out = {"t": t, "Y": Y, "R": R, "jitter": 1e-5, "observe_data": observe, "y_pseudo": yp,
               "kernel": onp.array([k[0] for k in kerns]), "hyper": onp.array([k[1:] for k in kerns]), "n_res": len(desc),
               "lml": float(lml), "mf": onp.asarray(res["m"]), "Pf": onp.asarray(res["P"]),
               "ms": onp.asarray(mu), "Ps": onp.asarray(var)}
        for p, dsc in enumerate(desc):
            out["w%d" % p] = onp.array(dsc["w"])
            out["term_kind%d" % p] = onp.array([k for k, _, _ in dsc["terms"]], dtype="U8")
            out["term_idx%d" % p] = onp.array([i for _, i, _ in dsc["terms"]], dtype=int)
            out["term_coef%d" % p] = onp.array([c for _, _, c in dsc["terms"]], dtype=float)
            out["forcing%d" % p] = onp.zeros(0)
        if bnd is not None:
            out["boundary"] = bnd[:, :, 0]
        fn = os.path.join(HERE, "ekfsys_%s_jit1e-5.npz" % name)
        onp.savez_compressed(fn, **out)
        written.append(fn)
    # =============================================== integrated Wiener prior (a6): the reference's own
    # WienerVelocity.{to_ss, expm, Q} (kernels/wiener.py:90-149) under the reference's sequential filter / smoother
    import scipy.special as ssp
    ns_w = {"np": jnp, "jax": jax, "chex": chex, "factorial": lambda n: Arr.__array_wrap__(onp.asarray(0.0), ssp.gamma(onp.asarray(n, float) + 1.0))
            if False else A(ssp.gamma(onp.asarray(n, dtype=float) + 1.0))}
    wv = mg.extract("kernels/wiener.py", ["WienerVelocity.to_ss", "WienerVelocity.expm", "WienerVelocity.Q",
                                           "WienerVelocity.state_size"], ns_w)

    class IWPPrior(sdes_mod.LTI_SDE):
        def __init__(self, q, var, ssc):
            self.k = types.SimpleNamespace(q=q, variance_param=types.SimpleNamespace(value=var), _state_space_dim=q + 1,
                                           stable_state_covariance=ssc, m_init=onp.zeros([q + 1, 1]).view(Arr))
            self.k.state_size = lambda: q + 1

        def _ss(self):
            return wv["WienerVelocity.to_ss"](self.k)

        def m_inf(self, x, X_s, t):
            return A(self._ss()[4])

        def P_inf(self, x, X_s, t):
            return A(self._ss()[5])

        def H(self, x, X_s, t):
            return A(self._ss()[3])

        def expm(self, X_s, dt):
            return A(wv["WienerVelocity.expm"](self.k, dt))

        def Q(self, dt, A_k, P_inf, X_spatial=None):
            return A(wv["WienerVelocity.Q"](self.k, dt, A_k, P_inf))
    for jitter in (1e-5,):
        settings.jitter = jitter
        for q_, var_, ssc_ in ((1, 0.7, 0.5), (2, 1.3, 0.2), (3, 0.4, 1.0)):
            rng = onp.random.default_rng(90 + q_)
            T = 50
            t = onp.cumsum(rng.uniform(0.5, 1.5, T) * 0.1)
            Y = onp.cumsum(rng.normal(size=(T, 1)) * 0.3, axis=0) + 0.1 * rng.normal(size=(T, 1))
            Y[rng.uniform(size=Y.shape) < 0.15] = onp.nan
            R = onp.tile(0.05 * onp.eye(1), [T, 1, 1])
            prior = IWPPrior(q_, var_, ssc_)
            data = types.SimpleNamespace(X_time=A(t), X_space=None, Nt=T, Ns=1, P=1, Y_st=A(Y[:, :, None]))
            lml, res = kf.filter_loop(data, prior, R=A(R), filter_type="sequential")
            out = {"t": t, "Y": Y, "R": R, "jitter": jitter, "q": q_, "variance": var_, "stable_state_covariance": ssc_,
                   "A_dt": onp.stack([onp.asarray(prior.expm(None, x)) for x in (0.0, 0.05, 0.9)]),
                   "Q_dt": onp.stack([onp.asarray(prior.Q(x, None, None)) for x in (0.0, 0.05, 0.9)]),
                   "lml": float(lml), "mf": onp.asarray(res["m"]), "Pf": onp.asarray(res["P"])}
            for fs in (False, True):
                mu, var = rts.smoother_loop(data, prior, res, full_state=fs, filter_type="sequential")
                out["ms_full%d" % fs], out["Ps_full%d" % fs] = onp.asarray(mu), onp.asarray(var)
            fn = os.path.join(HERE, "iwp_q%d.npz" % q_)
            onp.savez_compressed(fn, **out)
            written.append(fn)
    for f in written:
        print("wrote", os.path.relpath(f, HERE), os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
