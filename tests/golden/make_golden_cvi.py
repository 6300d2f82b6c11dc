#!/usr/bin/env python
"""Second golden generator: pins the CVI ASSEMBLY rows of SURVEY section 8 (a10, a12, a13) and the prior
stacking (a6) to the reference's own source, executed in place on the numpy stand-in of make_golden.py.

What runs from /root/reference (read in place, never copied):
  * computation/natural_gradients/cvi_nat_grad.py          `natural_gradients` (FullConjugateGaussian), ast-extracted,
        with the real `cvi_block_update`, the real `_get_fp_params` (cvi_nat_grad_utils.py:45-69) and the real
        theta_to_lambda / lambda_to_theta (exponential_family_transforms.py);
  * computation/elbos/elbos.py                              `elbo` (FullConjugateGaussian), ast-extracted;
  * computation/elbos/expected_log_likelihoods.py           `full_gaussian_expected_log_likelihood` (the leaf ELL);
  * computation/filters/{kalman_filter,rts_smoother}.py     the surrogate's posterior_blocks (sde_gp.py:255-277 glue);
  * computation/natural_gradients/cvi_hessian_approximations.py   the Gauss-Newton assembly of
        `_f_conditional_samples` (mask, J^T (-Lambda^-1) J, sums: the source lines between the "clean up shapes" and
        "return G" markers, executed as a slice) and `gauss_newton`'s 0.5 factor;
  * transforms/pdes.py                                      `DampedPendulum1D.forward` (the collocation residual);
  * transforms/transform.py                                 `Independent.{expm, P_inf, m_inf, H, Q}` with
        computation/matrix_ops.py `to_block_diag` / `get_block_diagonal` (batchjax.batch_or_loop in its documented
        loop mode).

What the stand-in supplies instead of JAX: `jax.jacfwd` by complex-step differentiation (exact to round-off for
the analytic residual), `jax.grad` by central differences with a large step (the Gaussian ELL is quadratic in
q_mu and linear in q_var, so central differences have NO truncation error), model / data / likelihood objects as
bare namespaces carrying the attributes the extracted functions read.  The evoke('marginal') /
evoke('expected_log_likelihood') dispatch for a Gaussian likelihood on the identity transform is replaced by the
leaf it resolves to (dispatched_ell.py:47-132 -> full_gaussian_expected_log_likelihood per time block).

    python tests/golden/make_golden_cvi.py        (needs /root/reference; writes tests/golden/cvi_assembly.npz,
                                                   gn_pendulum.npz, independent_stack.npz)
"""
import ast
import importlib
import os
import sys
import textwrap
import types

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

REF = mg.REF
Arr = mg.Arr


def A(x):
    return onp.asarray(x, dtype=onp.float64).view(Arr)


# ------------------------------------------------------------------------------ autodiff stand-ins
def jacfwd_complex(f, argnums=0):
    """d f / d arg by complex-step differentiation: Im f(x + i h e_j) / h with h = 1e-30."""
    idx = argnums[0] if isinstance(argnums, (list, tuple)) else argnums

    def g(*args):
        x = onp.asarray(args[idx], dtype=onp.float64)
        cols = []
        for j in range(x.size):
            xc = x.astype(onp.complex128).reshape(-1)
            xc[j] += 1e-30j
            a = list(args)
            a[idx] = xc.reshape(x.shape)
            cols.append(onp.imag(onp.asarray(f(*a))) / 1e-30)
        J = onp.stack(cols, -1).reshape(cols[0].shape + x.shape)
        return [J.view(Arr)] if isinstance(argnums, (list, tuple)) else J.view(Arr)
    return g


def grad_central(f, argnums, h=1e-2):
    """jax.grad of a scalar function w.r.t. array arguments by central differences."""
    def g(*args):
        outs = []
        for idx in argnums:
            x = onp.asarray(args[idx], dtype=onp.float64)
            gx = onp.zeros_like(x)
            flat = gx.reshape(-1)
            for j in range(x.size):
                e = onp.zeros(x.size)
                e[j] = h
                a, b = list(args), list(args)
                a[idx] = A(x + e.reshape(x.shape))
                b[idx] = A(x - e.reshape(x.shape))
                flat[j] = (float(f(*a)) - float(f(*b))) / (2 * h)
            outs.append(A(gx))
        return tuple(outs)
    return g


def extract_at(path, name, decorator_has, namespace):
    """Like make_golden.extract, but picks the overload of `name` whose @dispatch decorator mentions every string
    in `decorator_has` (the reference registers several functions of the same name under different keys)."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            decs = [ast.get_source_segment(src, d) or "" for d in node.decorator_list]
            if any(all(h in d for h in decorator_has) for d in decs):
                lineno = node.lineno
                node.decorator_list = []
                for a in node.args.args:                       # type annotations name classes that are not loaded
                    a.annotation = None
                node.returns = None
                ns = dict(namespace)
                exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, path), "exec"), ns)
                return ns[name], lineno
    raise KeyError((path, name, decorator_has))


def source_slice(path, start_marker, end_marker):
    """The reference's source lines from the first line containing start_marker (after any earlier occurrence
    of `after`) up to and including the first later line containing end_marker, dedented."""
    lines = open(os.path.join(REF, path)).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if start_marker in l)
    i1 = next(i for i in range(i0, len(lines)) if end_marker in lines[i])
    return textwrap.dedent("\n".join(lines[i0:i1 + 1])), (i0 + 1, i1 + 1)


def main():
    assert os.path.isdir(REF), "the reference tree is needed to (re)generate the golden vectors"
    jax = mg.install_standin()
    sdes_mod = mg.install_package_tree()
    settings = sys.modules["stgp.settings"]
    settings.verbose = False
    jnp = jax.numpy
    # jax accepts a LIST of axes in reductions; numpy wants a tuple
    jnp._extra["sum"] = lambda a, axis=None, **k: onp.sum(a, axis=tuple(axis) if isinstance(axis, list) else axis,
                                                          **k).view(Arr)
    chex = sys.modules["chex"]
    kf = importlib.import_module("stgp.computation.filters.kalman_filter")
    rts = importlib.import_module("stgp.computation.filters.rts_smoother")
    eft = importlib.import_module("stgp.computation.natural_gradients.exponential_family_transforms")
    mo = importlib.import_module("stgp.computation.matrix_ops")
    nu = importlib.import_module("stgp.utils.nan_utils")
    ga = importlib.import_module("stgp.computation.gaussian")
    disp = importlib.import_module("stgp.dispatch")
    Kern, Prior = mg.build_prior_classes(jax, sdes_mod)
    partial = __import__("functools").partial
    written = []

    base_ns = {"np": jnp, "chex": chex, "jit": mg._jit, "jax": jax, "settings": settings, "partial": partial,
               "_ensure_str": disp._ensure_str}
    for k in ("cholesky", "cholesky_solve", "add_jitter", "to_block_diag", "get_block_diagonal"):
        base_ns[k] = getattr(mo, k)
    for k in ("get_mask", "mask_to_identity", "mask_vector", "get_same_shape_mask"):
        base_ns[k] = getattr(nu, k)
    for k in ("log_gaussian", "log_gaussian_with_nans", "log_gaussian_scalar", "log_gaussian_with_mask"):
        if hasattr(ga, k):
            base_ns[k] = getattr(ga, k)
    block_update = mg.extract("computation/natural_gradients/cvi_nat_grad.py", ["cvi_block_update"],
                              base_ns)["cvi_block_update"]
    leaf_ell = mg.extract("computation/elbos/expected_log_likelihoods.py", ["full_gaussian_expected_log_likelihood"],
                          base_ns)["full_gaussian_expected_log_likelihood"]

    # =============================================================== natural_gradients + elbo, Gaussian likelihood
    out = {}
    for case, (latents, T, seed, nan_frac) in {
            "m32_fs": ([[("m32", 0.9, 1.2)]], 14, 3, 0.2),
            "m52_fs": ([[("m52", 0.7, 0.8)]], 10, 4, 0.0)}.items():
        for ngj in (1e-7, 1e-5):
            settings.jitter, settings.ng_jitter = 1e-5, ngj
            rng = onp.random.default_rng(seed)
            prior = Prior([[Kern(*k) for k in lat] for lat in latents], True)     # full-state sites: D = d
            D = prior.P_inf(None, None, None).shape[0]
            t = onp.cumsum(rng.uniform(0.5, 1.5, T) * 0.1)
            Ytil = 0.5 * rng.normal(size=(T, D))
            G = rng.normal(size=(T, D, D)) * 0.3
            Vtil = G @ onp.swapaxes(G, -1, -2) + 0.5 * onp.eye(D)
            # the model's data likelihood: Gaussian noise on the (identity-transformed) site block
            Yobs = 0.7 * rng.normal(size=(T, D))
            Yobs[rng.uniform(size=Yobs.shape) < nan_frac] = onp.nan
            Gn = rng.normal(size=(D, D)) * 0.2
            noise = Gn @ Gn.T + 0.3 * onp.eye(D)

            # --- q.surrogate: an SDE_GP whose data / noise are the sites (sde_gp.py:255-277 posterior_blocks)
            sur_data = types.SimpleNamespace(X_time=A(t), X_space=None, Nt=T, Ns=1, P=D, Y_st=A(Ytil[:, :, None]),
                                             _Y=types.SimpleNamespace(value=A(Ytil)), minibatch=False, N=T)

            def posterior_blocks(return_lml=False):
                lml, res = kf.filter_loop(sur_data, prior, R=A(Vtil), filter_type="sequential")
                mu, var = rts.smoother_loop(sur_data, prior, res, full_state=False, filter_type="sequential")
                mu, var = A(mu), A(onp.asarray(var)[:, None, ...])
                return (lml, mu, var) if return_lml else (mu, var)
            sur_lik = types.SimpleNamespace(variance=A(Vtil))
            surrogate = types.SimpleNamespace(posterior_blocks=posterior_blocks, data=sur_data, likelihood=sur_lik)
            q = types.SimpleNamespace(surrogate=surrogate)
            base_prior = types.SimpleNamespace(get_sparsity_list=lambda: None)
            model_prior = types.SimpleNamespace(base_prior=base_prior)
            data = types.SimpleNamespace(minibatch=False, N=T, Y=A(Yobs))
            lik = types.SimpleNamespace(noise=A(noise))
            model = types.SimpleNamespace(approximate_posterior=q, prior=model_prior, data=data, likelihood=lik,
                                          inference=types.SimpleNamespace(whiten=False))

            def ell_with_variational_params(dat, q_m, q_S, likelihood, pr, ap, inference):
                """elbos.py:19-43 with the marginal / ELL dispatch resolved for a Gaussian likelihood on the
                identity transform: per time block full_gaussian_expected_log_likelihood, summed."""
                if likelihood is sur_lik:
                    Ys, Ns = onp.asarray(sur_data._Y.value), onp.asarray(sur_lik.variance)
                else:
                    Ys, Ns = onp.asarray(dat.Y), onp.broadcast_to(onp.asarray(likelihood.noise), (T, D, D))
                tot = 0.0
                for k in range(T):
                    tot = tot + leaf_ell(A(onp.zeros([D, 1])), A(Ys[k][:, None]), A(Ns[k]), A(onp.asarray(q_m)[k]),
                                         A(onp.asarray(q_S)[k, 0]))
                return tot

            def partial_ell(m, q_m, q_S):                      # cvi_nat_grad_utils.py:156-168
                return ell_with_variational_params(m.data, q_m, q_S, m.likelihood, m.prior, m.approximate_posterior,
                                                   m.inference)

            ns = dict(base_ns)
            ns["theta_to_lambda"] = eft.theta_to_lambda
            ns["theta_precision_to_lambda"] = getattr(eft, "theta_precision_to_lambda", None)
            ns["print"] = lambda *a, **k: None
            fp = mg.extract("computation/natural_gradients/cvi_nat_grad_utils.py", ["_get_fp_params"], ns)
            ns["_get_fp_params"] = fp["_get_fp_params"]
            ns["partial_ell"] = partial_ell
            ns["cvi_block_update"] = block_update
            ns["GAUSS_NEWTON_ENFORCE_TYPES"] = []
            jax_ng = types.SimpleNamespace(**{k: getattr(jax, k) for k in ("vmap", "jit", "numpy", "lax")})
            jax_ng.grad = lambda f, argnums: grad_central(f, argnums)
            ns["jax"] = jax_ng
            ng, ng_line = extract_at("computation/natural_gradients/cvi_nat_grad.py", "natural_gradients",
                                     ["FullConjugateGaussian", "NoSparsity"], ns)
            beta = 0.4
            l1n, l2n = ng(model, beta, None, "NG_Moment")
            # NG_Moment re-entry (cvi_parameterisations.py:63-93): lambda -> theta per block
            th = [eft.lambda_to_theta(A(onp.asarray(l1n)[k]), A(onp.asarray(l2n)[k])) for k in range(T)]
            Yn = onp.stack([onp.asarray(a)[:, 0] for a, _ in th])
            Vn = onp.stack([onp.asarray(b) for _, b in th])

            ns_e = dict(base_ns)
            ns_e["compute_expected_log_liklihood_with_variational_params"] = ell_with_variational_params
            ns_e["print"] = lambda *a, **k: None
            elbo_fn, elbo_line = extract_at("computation/elbos/elbos.py", "elbo", ["FullConjugateGaussian"], ns_e)
            elbo_val = elbo_fn(data, lik, model_prior, q, model.inference)
            lml, q_m, q_S = posterior_blocks(True)
            key = "%s_ngj%s" % (case, "1e-7" if ngj == 1e-7 else "1e-5")
            for nm, val in (("t", t), ("Ytil", Ytil), ("Vtil", Vtil), ("Yobs", Yobs), ("noise", noise), ("beta", beta),
                            ("ng_jitter", ngj), ("jitter", 1e-5), ("P_inf", onp.asarray(prior.P_inf(None, None, None))),
                            ("q_mu", onp.asarray(q_m)), ("q_var", onp.asarray(q_S)), ("lml", float(lml)),
                            ("lambda1_new", onp.asarray(l1n)), ("lambda2_new", onp.asarray(l2n)),
                            ("Ytil_new", Yn), ("Vtil_new", Vn), ("elbo", float(elbo_val))):
                out["%s_%s" % (key, nm)] = onp.asarray(val)
            out["%s_kernel" % key] = onp.array([latents[0][0][0]])
            out["%s_hyper" % key] = onp.array(latents[0][0][1:])
    fn = os.path.join(HERE, "cvi_assembly.npz")
    onp.savez_compressed(fn, **out)
    written.append(fn)

    # ======================================================= Gauss-Newton curvature of the damped-oscillator model
    fwd = mg.extract("transforms/pdes.py", ["DampedPendulum1D.forward"], {"np": jnp})["DampedPendulum1D.forward"]
    gn_src, gn_lines = source_slice("computation/natural_gradients/cvi_hessian_approximations.py",
                                    "if _ensure_str(data) == 'TemporallyGroupedData':\n        Y_st_mask", "G = G[:, None, ...]") \
        if False else (None, None)
    # the slice starts at the mask construction that follows the "clean up shapes" block of _f_conditional_samples
    lines = open(os.path.join(REF, "computation/natural_gradients/cvi_hessian_approximations.py")).read().split("\n")
    i_clean = next(i for i, l in enumerate(lines) if "# clean up shapes" in l)
    i0 = next(i for i in range(i_clean, len(lines)) if "_ensure_str(data) == 'TemporallyGroupedData'" in lines[i])
    i1 = next(i for i in range(i0, len(lines)) if "G = G[:, None, ...]" in lines[i])
    gn_src = textwrap.dedent("\n".join(lines[i0:i1 + 1]))
    rng = onp.random.default_rng(21)
    T, Dd = 9, 4
    g_, l_, b_ = 9.81, 1.3, 0.35
    var_obs, var_col = 0.05 ** 2, 0.4 ** 2
    pend = types.SimpleNamespace(g_param=types.SimpleNamespace(value=g_), l_param=types.SimpleNamespace(value=l_),
                                 b_param=types.SimpleNamespace(value=b_))
    u = rng.normal(size=(T, Dd)) * onp.array([1.0, 2.0, 4.0, 8.0])

    def T_of_u(uu):
        """MultiOutput([observe x, DampedPendulum1D residual]) of zoo/sde_diff.py:757-763 (restated stacking):
        output 0 = x, output 1 = reference forward(x, x_t, x_tt)."""
        return jnp.array([uu[0], fwd(pend, uu[:3])[0]])
    J = onp.stack([onp.asarray(jacfwd_complex(T_of_u)(u[k])) for k in range(T)])          # [T, P=2, D]
    Y = onp.stack([rng.normal(size=T), onp.zeros(T)], -1)                                   # [T, P]
    Y[rng.uniform(size=T) < 0.4, 0] = onp.nan
    Y[3, 1] = onp.nan
    # shapes the slice expects (single spatial point): J_u_tf [Nt, Ns, P, Ms, 1], neg_Lambda [Nt, Ns, P, 1, 1],
    # Y_st [Nt, P, Ns]
    env = dict(base_ns)

    class TemporalData:                                          # dispatch key only
        minibatch = False
    env.update(J_u_tf=A(J[:, None, :, :, None]),
               neg_Lambda=A(onp.broadcast_to(-1.0 / onp.array([var_obs, var_col]), (T, 1, 2))[..., None, None].copy()),
               Y_st=A(Y[:, :, None]), data=TemporalData())
    exec(compile(gn_src, "cvi_hessian_approximations.py[%d:%d]" % (i0 + 1, i1 + 1), "exec"), env)
    G_ref = onp.asarray(env["G"])                                 # [Nt, 1, Ms, Ms]
    approx_hessian = 0.5 * G_ref                                  # gauss_newton(): approx_hessian = 0.5 * G
    fn = os.path.join(HERE, "gn_pendulum.npz")
    onp.savez_compressed(fn, u=u, Y=Y, J=J, g=g_, l=l_, b=b_, var_obs=var_obs, var_col=var_col,
                         approx_hessian=approx_hessian[:, 0], residual=onp.stack(
                             [onp.asarray(fwd(pend, u[k][:3]))[0] for k in range(T)]),
                         slice_lines=onp.array([i0 + 1, i1 + 1]))
    written.append(fn)

    # ============================================================== Independent stacking (transform.py:400-545)
    def batch_or_loop(fn_, inputs, axes, dim, out_dim, batch_type=None):
        """batchjax.batch_or_loop in loop mode (utils/utils.py:60-77 falls back to it for mixed latent types)."""
        res = []
        for i in range(dim):
            args = [inp if ax is None else inp[i] for inp, ax in zip(inputs, axes)]
            res.append(fn_(*args))
        if out_dim == 1:
            return res
        return [[r[j] for r in res] for j in range(out_dim)]
    ns_i = dict(base_ns)
    ns_i.update(batch_or_loop=batch_or_loop, get_batch_type=lambda p: "loop", warnings=__import__("warnings"))
    ind = mg.extract("transforms/transform.py", ["Independent.expm", "Independent.P_inf", "Independent.m_inf",
                                                 "Independent.H", "Independent.Q",
                                                 "Independent.state_space_representation"], ns_i)
    markov_Q = mg.extract("kernels/kernel.py", ["MarkovKernel.Q"], base_ns)["MarkovKernel.Q"]

    class KWrap:
        """kernel facade with the MarkovKernel API (kernel.py:163-209) over make_golden's per-kernel closed forms"""
        def __init__(self, k):
            self.k = k

        def expm(self, dt, x_s):
            return self.k.expm(dt)

        def to_ss(self, x_s):
            return self.k.to_ss()

        def P_inf(self, x, X_s, t):
            return self.k.to_ss()[5]

        def m_inf(self, x, X_s, t):
            return self.k.to_ss()[4]

        def H(self, x, X_s, t):
            return self.k.to_ss()[3]

        def Q(self, dt, A_k, P, X_spatial=None):
            return markov_Q(None, dt, A_k, P)
    st = {}
    for name, kinds in {"m32_m32": [("m32", 1.0, 1.3), ("m32", 0.4, 0.5)], "m52x3": [("m52", 0.7, 0.9)] * 3}.items():
        ks = [Kern(*k) for k in kinds]
        sdim = onp.asarray(ks[0].to_ss()[5]).shape[0]
        parent = [types.SimpleNamespace(kernel=KWrap(k)) for k in ks]
        self_ = types.SimpleNamespace(parent=parent, output_dim=len(ks), state_space_dim=lambda: [sdim] * len(ks),
                                      spatial_output_dim=[1] * len(ks))
        for dt in (0.05, 0.9):
            A_k = onp.asarray(ind["Independent.expm"](self_, dt, None))
            P_inf = onp.asarray(ind["Independent.P_inf"](self_, None, None, None))
            Qk = onp.asarray(ind["Independent.Q"](self_, dt, A(A_k), A(P_inf), None))
            key = "%s_dt%s" % (name, dt)
            st[key + "_A"], st[key + "_Pinf"], st[key + "_Q"] = A_k, P_inf, Qk
        # what LTI_SDE.{H, m_inf, P_inf} hand to the filter (sdes.py:58-90): the state_space_representation tuple
        F_, L_, Qc_, H_, minf_, Pinf_ = ind["Independent.state_space_representation"](self_, None)
        st[name + "_H"], st[name + "_minf"] = onp.asarray(H_), onp.asarray(minf_)
        st[name + "_F"], st[name + "_Pinf_ssr"] = onp.asarray(F_), onp.asarray(Pinf_)
        # the accessor methods of the same class stack differently (hstack): recorded, unused by the filter
        st[name + "_H_accessor"] = onp.asarray(ind["Independent.H"](self_, None, None, None))
        st[name + "_kinds"] = onp.array([k[0] for k in kinds])
        st[name + "_hyper"] = onp.array([k[1:] for k in kinds])
    fn = os.path.join(HERE, "independent_stack.npz")
    onp.savez_compressed(fn, **st)
    written.append(fn)
    # ===================================== Monte-Carlo ELL of the reference (the path config 4 really takes)
    # mv_indepentdent_monte_carlo (integrals/approximators.py:16-58) with the reference's own scalar
    # log-likelihoods (general.py:9-26, likelihood/poisson.py:20-22, likelihood/bernoulli.py:17-19), driven by a
    # numpy PRNG in place of objax.random (the objax threefry stream is not reproducible here): the fixture records
    # the MC mean and its standard error, which tests/test_oracle_cvi.py uses as a STATISTICAL pin (4 sigma) for
    # the Gauss-Hermite quadrature that replaces MC on the B200 path.
    import scipy.special as ssp
    ns_m = dict(base_ns)
    mc_rng = onp.random.default_rng(2024)
    ns_m["objax"] = types.SimpleNamespace(random=types.SimpleNamespace(
        normal=lambda shape, mean=0.0, stddev=1.0, generator=None: A(mc_rng.normal(size=shape))))
    ns_m["jax"] = types.SimpleNamespace(vmap=lambda f, in_axes, out_axes=0: (
        # the generic stand-in vmap loops in Python; here only argument 1 (the samples) is batched and the
        # reparameterisation is elementwise, so the batch is evaluated sample-by-sample in chunks
        lambda fn, samples, mu, var, *args: onp.stack([onp.asarray(f(fn, A(s), mu, var, *args)) for s in samples])))
    mc = mg.extract("computation/integrals/approximators.py", ["mv_indepentdent_monte_carlo"], ns_m)[
        "mv_indepentdent_monte_carlo"]
    gen_ns = {"np": jnp, "jit": mg._jit, "chex": types.SimpleNamespace(assert_rank=lambda *a: None),
              "gammaln": lambda x: A(ssp.gammaln(x))}
    logs = mg.extract("computation/general.py", ["log_poisson", "log_bernoulli"], gen_ns)
    S = 200000
    sites = [(3.0, 0.4, 0.3), (0.0, -0.5, 0.8), (7.0, 1.2, 0.15), (1.0, 0.1, 1.5)]
    mcout = {"num_samples": S}
    for kind in ("poisson", "bernoulli"):
        ys = onp.array([s[0] if kind == "poisson" else float(s[0] > 1) for s in sites])
        mu = A(onp.array([s[1] for s in sites])[:, None, None])
        var = A(onp.array([s[2] for s in sites])[:, None, None])
        binsize = 0.8
        if kind == "poisson":
            fn_ = lambda f: A(logs["log_poisson"](ys[:, None, None], jnp.exp(f) * binsize))   # noqa: E731
        else:
            fn_ = lambda f: A(ys[:, None, None] * jnp.log(A(ssp.ndtr(f)) + 1e-5)              # noqa: E731
                              + (1 - ys[:, None, None]) * jnp.log(1 - A(ssp.ndtr(f)) + 1e-5))
        samp = onp.asarray(mc(fn_, mu, var, generator=object(), num_samples=S, average=False))   # [S, N, 1, 1]
        mcout[kind + "_y"], mcout[kind + "_m"], mcout[kind + "_v"] = ys, onp.asarray(mu)[:, 0, 0], onp.asarray(var)[:, 0, 0]
        mcout[kind + "_mc_mean"] = samp.mean(0)[:, 0, 0]
        mcout[kind + "_mc_stderr"] = samp.std(0, ddof=1)[:, 0, 0] / onp.sqrt(S)
        mcout[kind + "_binsize"] = binsize
    fn = os.path.join(HERE, "mc_ell.npz")
    onp.savez_compressed(fn, **mcout)
    written.append(fn)
    # ================================================== data sort / pad: the reference's own numpy
    # helpers pad_with_nan_to_make_grid and order_sequentially_np (data/sequential.py:9-144), as SequentialData.sort /
    # unsort compose them (data/data.py:353-415)
    ns_d = {"onp": onp, "np": jnp, "chex": chex}
    seq = mg.extract("data/sequential.py", ["pad_with_nan_to_make_grid", "order_sequentially_np"], ns_d)
    rng = onp.random.default_rng(77)
    tt_, ss_ = onp.round(onp.sort(rng.uniform(0, 3, 7)), 3), onp.round(rng.uniform(0, 1, (4, 2)), 3)
    full = onp.array([[a, *b] for a in tt_ for b in ss_])
    keep = rng.permutation(full.shape[0])[:19]                      # scattered subset, shuffled
    X = onp.vstack([full[keep], full[keep[:3]]])                    # with three duplicated locations
    Y = rng.normal(size=(X.shape[0], 2))
    added, Xp, Yp = seq["pad_with_nan_to_make_grid"](X, Y)
    uidx, ridx, sidx, Xs, Ys = seq["order_sequentially_np"](A(Xp), A(Yp))
    Yst = onp.transpose(onp.asarray(Ys), [0, 2, 1])
    payload = rng.normal(size=(onp.asarray(Xs).shape[0] * onp.asarray(Xs).shape[1], 3))
    unsorted = payload[onp.asarray(sidx)][onp.asarray(ridx)][:X.shape[0]]
    fn = os.path.join(HERE, "sort_pad.npz")
    onp.savez_compressed(fn, X=X, Y=Y, points_added=added, X_padded=onp.asarray(Xp), Y_padded=onp.asarray(Yp),
                         unique_idx=onp.asarray(uidx), reverse_idx=onp.asarray(ridx), sort_idx=onp.asarray(sidx),
                         X_sorted=onp.asarray(Xs), Y_st=Yst, payload=payload, unsorted=unsorted)
    written.append(fn)
    for f in written:
        print("wrote", os.path.relpath(f, HERE), os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
