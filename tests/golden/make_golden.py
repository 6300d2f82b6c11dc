#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES (read in place from
/root/reference, never copied) on a numpy-backed stand-in for jax.

Why a stand-in: jax / jaxlib / objax / chex / batchjax are not installable in the build container (no
network), so `import stgp` is impossible.  The hot-path files, however, only use a small, purely
functional slice of the jax API (jax.numpy array algebra, lax.scan, lax.associative_scan, vmap, jit,
jax.scipy.linalg).  This script installs minimal modules of those names backed by numpy / scipy
(LAPACK fp64, the same factorisations XLA's CPU backend calls), registers empty stand-ins for the
package __init__ files (so that importing e.g. stgp.computation.filters.kalman_filter executes THAT
file and its real dependencies matrix_ops.py, linalg.py, gaussian.py, nan_utils.py, dispatch.py,
settings.py -- but not the whole model zoo), and then calls the reference functions through the
reference's own dispatch registry: evoke('filter', 'sequential'), evoke('smoother', 'sequential'),
evoke('filter', 'parallel'), evoke('smoother', 'parallel'), theta_to_lambda / lambda_to_theta.
Functions that live in modules with heavy import chains (cvi_block_update, the closed-form block ELL,
the Matern-5/2 / 7/2 `expm` / `to_ss` methods) are extracted from the reference file by name with
`ast` and compiled unchanged against the same stand-in.

The inputs and the reference's outputs are stored as .npz; tests/test_golden.py pins the oracle (and
tests/test_gpu_golden.py the CUDA path) against them.  /root/reference is only needed to (re)generate.

    python tests/golden/make_golden.py
"""
import ast
import importlib
import os
import sys
import types

import numpy as onp
import scipy.linalg as sla

REF = "/root/reference/src/lib/stgp"
OUT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------- jax stand-in
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, v):
                out = onp.array(arr, copy=True)
                out[idx] = v
                return out.view(Arr)

            def add(self, v):
                out = onp.array(arr, copy=True)
                out[idx] += v
                return out.view(Arr)
        return _Setter()


class Arr(onp.ndarray):
    """ndarray with jax's functional `.at[idx].set(v)`."""
    @property
    def at(self):
        return _At(self)


def _wrap_out(x):
    if isinstance(x, onp.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, tuple):
        return tuple(_wrap_out(v) for v in x)
    if isinstance(x, list):
        return [_wrap_out(v) for v in x]
    return x


def _wrap_fn(f):
    def g(*a, **k):
        return _wrap_out(f(*a, **k))
    g.__name__ = getattr(f, "__name__", "f")
    return g


class _Proxy(types.ModuleType):
    """Module whose attributes are those of `backend`, with ndarray results viewed as Arr."""
    def __init__(self, name, backend, extra=None):
        super().__init__(name)
        self.__dict__["_backend"] = backend
        self.__dict__["_extra"] = extra or {}

    def __getattr__(self, name):
        if name in self._extra:
            return self._extra[name]
        v = getattr(self._backend, name)
        if isinstance(v, type) or not callable(v):
            return v
        return _wrap_fn(v)


def _tree_map(f, *trees):
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: _tree_map(f, *[t[k] for t in trees]) for k in t0}
    if isinstance(t0, (tuple, list)):
        return type(t0)(_tree_map(f, *[t[i] for t in trees]) for i in range(len(t0)))
    if t0 is None:
        return None
    return f(*trees)


def _tree_leaves(t):
    if isinstance(t, dict):
        return [l for k in t for l in _tree_leaves(t[k])]
    if isinstance(t, (tuple, list)):
        return [l for v in t for l in _tree_leaves(v)]
    return [] if t is None else [t]


def _jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def _vmap(f, in_axes=0, out_axes=0):
    def g(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else [in_axes] * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _tree_leaves(a)[0].shape[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else _tree_map(lambda x: onp.take(onp.asarray(x), i, axis=ax).view(Arr), a)
                  for a, ax in zip(args, axes)]
            outs.append(f(*sl))
        return _tree_map(lambda *xs: onp.stack([onp.asarray(x) for x in xs], axis=out_axes).view(Arr), *outs)
    return g


def _scan(f, init, xs, length=None, reverse=False, unroll=1):
    n = _tree_leaves(xs)[0].shape[0] if xs is not None else length
    carry, ys = init, []
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        x = None if xs is None else _tree_map(lambda a: onp.asarray(a)[i].view(Arr) if onp.ndim(a[i]) else a[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    if reverse:
        ys = ys[::-1]
    return carry, _tree_map(lambda *v: onp.stack([onp.asarray(u) for u in v]).view(Arr), *ys)


def _associative_scan(fn, elems, reverse=False, axis=0):
    """The odd/even recursion of jax.lax.associative_scan (jax/_src/lax/control_flow/loops.py), so the
    combine order -- and therefore the floating-point result -- is the one jax produces."""
    assert axis == 0
    flat = list(elems)
    if reverse:
        flat = [onp.flip(onp.asarray(e), 0) for e in flat]
    flat = [onp.asarray(e) for e in flat]

    def combine(a, b):
        if a[0].shape[0] == 0:                      # jax evaluates the vmapped operator on an empty batch
            return [x[:0] for x in a]
        return [onp.asarray(x) for x in fn(tuple(x.view(Arr) for x in a), tuple(x.view(Arr) for x in b))]

    def interleave(a, b):
        out = onp.empty((a.shape[0] + b.shape[0],) + a.shape[1:], a.dtype)
        out[0::2] = a
        out[1::2] = b
        return out

    def rec(es):
        n = es[0].shape[0]
        if n < 2:
            return es
        reduced = combine([e[0:-1:2] for e in es], [e[1::2] for e in es])
        odd = rec(reduced)
        if n % 2 == 0:
            even = combine([e[:-1] for e in odd], [e[2::2] for e in es])
        else:
            even = combine(odd, [e[2::2] for e in es])
        even = [onp.concatenate([e[0:1], r], 0) for e, r in zip(es, even)]
        return [interleave(a, b) for a, b in zip(even, odd)]

    res = rec(flat)
    if reverse:
        res = [onp.flip(r, 0) for r in res]
    return tuple(r.view(Arr) for r in res)


def _not_available(name):
    def f(*a, **k):
        raise NotImplementedError("%s is not provided by the numpy stand-in" % name)
    return f


def _solve(a, b, assume_a="gen", **kw):
    return sla.solve(a, b, assume_a={"gen": "gen", "pos": "pos", "sym": "sym"}.get(assume_a, "gen"))


def install_standin():
    jnp_linalg = _Proxy("jax.numpy.linalg", onp.linalg)
    jnp = _Proxy("jax.numpy", onp, {"linalg": jnp_linalg, "ndarray": onp.ndarray,
                                      "vectorize": onp.vectorize})
    jsp_linalg = _Proxy("jax.scipy.linalg", sla, {"solve": _wrap_fn(_solve)})
    jsp_sparse_linalg = types.ModuleType("jax.scipy.sparse.linalg")
    jsp_sparse_linalg.cg = _not_available("cg")
    jsp_sparse = types.ModuleType("jax.scipy.sparse")
    jsp_sparse.linalg = jsp_sparse_linalg
    import scipy.special as ssp
    jsp_special = _Proxy("jax.scipy.special", ssp)
    jsp = types.ModuleType("jax.scipy")
    jsp.linalg, jsp.sparse, jsp.special = jsp_linalg, jsp_sparse, jsp_special
    lax = types.ModuleType("jax.lax")
    lax.scan, lax.associative_scan = _scan, _associative_scan
    lax.stop_gradient = lambda x: x
    jax = types.ModuleType("jax")
    jax.numpy, jax.scipy, jax.lax = jnp, jsp, lax
    jax.jit, jax.vmap = _jit, _vmap
    for n in ("jacfwd", "jacrev", "grad", "vjp", "jvp", "hessian", "value_and_grad"):
        setattr(jax, n, _not_available(n))
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.Array = type("Array", (), {})          # scipy's array-api probing looks these up once `jax` is importable
    jax.core = types.SimpleNamespace(Tracer=type("Tracer", (), {}))
    jax.__version__ = "0.0-numpy-standin"
    jax.tree_util = types.SimpleNamespace(tree_map=_tree_map)
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jnp_linalg, "jax.scipy": jsp,
            "jax.scipy.linalg": jsp_linalg, "jax.scipy.sparse": jsp_sparse,
            "jax.scipy.sparse.linalg": jsp_sparse_linalg, "jax.scipy.special": jsp_special, "jax.lax": lax}

    chex = types.ModuleType("chex")

    def assert_rank(x, r):
        xs = x if isinstance(x, (list, tuple)) else [x]
        rs = r if isinstance(r, (list, tuple)) else [r] * len(xs)
        for a, b in zip(xs, rs):
            assert onp.ndim(a) == b, "chex.assert_rank: %s != %s" % (onp.ndim(a), b)

    def assert_shape(x, s):
        if isinstance(x, (list, tuple)):
            for a, b in zip(x, s):
                assert_shape(a, b)
            return
        assert tuple(onp.shape(x)) == tuple(s), "chex.assert_shape: %s != %s" % (onp.shape(x), s)

    def assert_equal(a, b):
        assert a == b, "chex.assert_equal: %s != %s" % (a, b)
    chex.assert_rank, chex.assert_shape, chex.assert_equal = assert_rank, assert_shape, assert_equal
    chex.assert_equal_shape = lambda xs: None
    mods["chex"] = chex

    objax = types.ModuleType("objax")

    class Module:
        pass

    class ModuleList(list):
        pass
    objax.Module, objax.ModuleList = Module, ModuleList
    for n in ("TrainVar", "StateVar", "TrainRef", "Jit", "Vectorize", "VarCollection"):
        setattr(objax, n, type(n, (), {}))
    objax.random = types.SimpleNamespace(normal=_not_available("objax.random.normal"), Generator=object)
    objax.functional = types.SimpleNamespace()
    mods["objax"] = objax
    mods["bibtexparser"] = types.ModuleType("bibtexparser")
    sys.modules.update(mods)
    return jax


def install_package_tree():
    """`stgp` and its sub-packages as empty namespace modules whose __path__ points at the reference
    tree: sub-MODULES load from the real files, the heavyweight __init__.py files are never executed.
    transforms.sdes / transforms.pdes (deep objax class hierarchies) are replaced by bare classes of the
    same names; the filters only use them as dispatch keys."""
    def pkg(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, rel)]
        m.__package__ = name
        sys.modules[name] = m
        return m
    pkg("stgp", "")
    for sub in ("computation", "computation/filters", "computation/natural_gradients", "computation/elbos",
                "utils", "kernels", "transforms"):
        pkg("stgp." + sub.replace("/", "."), sub)
    sdes = types.ModuleType("stgp.transforms.sdes")

    class SDE:
        pass

    class LTI_SDE(SDE):
        pass

    class LinearizedFilter_SDE(SDE):
        pass
    sdes.SDE, sdes.LTI_SDE, sdes.LinearizedFilter_SDE = SDE, LTI_SDE, LinearizedFilter_SDE
    pdes = types.ModuleType("stgp.transforms.pdes")

    class PDE:
        pass
    pdes.PDE = PDE
    sys.modules["stgp.transforms.sdes"] = sdes
    sys.modules["stgp.transforms.pdes"] = pdes
    stgp = sys.modules["stgp"]
    stgp.settings = importlib.import_module("stgp.settings")
    stgp.dispatch = importlib.import_module("stgp.dispatch")
    return sdes


def extract(path, names, namespace):
    """Compile the named top-level functions (or 'Class.method') of a reference file, unchanged, in
    `namespace`.  Decorators are kept (the namespace provides jit / partial / dispatch)."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    out = {}
    for name in names:
        cls, _, fn = name.rpartition(".")
        body = tree.body
        if cls:
            body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
        node = [n for n in body if isinstance(n, ast.FunctionDef) and n.name == fn][-1]
        node.decorator_list = [d for d in node.decorator_list
                               if not (isinstance(d, ast.Call) and getattr(d.func, "id", "") == "dispatch")]
        mod = ast.Module(body=[node], type_ignores=[])
        ns = dict(namespace)
        exec(compile(mod, os.path.join(REF, path), "exec"), ns)
        out[name] = ns[fn]
    return out


# ------------------------------------------------------------------------------------------- priors
def build_prior_classes(jax, sdes_mod):
    jnp = jax.numpy
    ss = importlib.import_module("stgp.kernels.ss_utils")
    ns = {"np": jnp, "chex": sys.modules["chex"], "jit": _jit, "jax": jax}
    mat = extract("kernels/matern.py", ["Matern52.to_ss", "Matern52.expm", "ScaledMatern72.to_ss",
                                        "ScaledMatern72.expm"], ns) \
        if _has_class("kernels/matern.py", "ScaledMatern72") else \
        extract("kernels/matern.py", ["Matern52.to_ss", "Matern52.expm", "Matern72.to_ss", "Matern72.expm"], ns)
    k72 = "ScaledMatern72" if "ScaledMatern72.to_ss" in mat else "Matern72"
    markov_Q = extract("kernels/kernel.py", ["MarkovKernel.Q"], ns)["MarkovKernel.Q"]

    class Kern:
        """One temporal Markov kernel evaluated by the reference's own closed forms."""
        def __init__(self, kind, ls, var):
            self.kind, self.ls, self.var = kind, float(ls), float(var)
            self.self_ = types.SimpleNamespace(variance=self.var, lengthscales=[self.ls], input_dim=1)

        def to_ss(self):
            if self.kind == "m32":
                return ss.matern32_temporal_state_space_rep(self.ls, self.var)      # ss_utils.py:12-38
            if self.kind == "m52":
                return mat["Matern52.to_ss"](self.self_)                             # matern.py:115-146
            return mat[k72 + ".to_ss"](self.self_)                                   # matern.py:275-301

        def expm(self, dt):
            if self.kind == "m32":
                return ss.matern32_temporal_expm(dt, self.ls)                        # ss_utils.py:6-10
            if self.kind == "m52":
                return mat["Matern52.expm"](self.self_, dt)                          # matern.py:152-177
            return mat[k72 + ".expm"](self.self_, dt)                                # matern.py:306-329

    class Prior(sdes_mod.LTI_SDE):
        """Duck-typed LTI_SDE over Independent / Sum stacks of `Kern`s.  The stacking (block-diagonal F,
        P_inf, A; H rows per latent: transform.py:400-545, kernel.py:134-160; full-state H = identity:
        sdes.py:99-172) is restated here; every per-kernel quantity comes from the reference code."""
        def __init__(self, latents, full_state_obs=False):
            self.latents, self.fso = latents, full_state_obs

        def _blocks(self, f):
            return sla.block_diag(*[onp.asarray(f(k)) for lat in self.latents for k in lat])

        def m_inf(self, x, X_s, t):
            return onp.vstack([onp.asarray(k.to_ss()[4]) for lat in self.latents for k in lat]).view(Arr)

        def P_inf(self, x, X_s, t):
            return self._blocks(lambda k: k.to_ss()[5]).view(Arr)

        def H(self, x, X_s, t):
            d = self.P_inf(None, None, None).shape[0]
            if self.fso:
                return onp.eye(d).view(Arr)
            rows = [onp.hstack([onp.asarray(k.to_ss()[3]) for k in lat]) for lat in self.latents]
            return sla.block_diag(*rows).view(Arr)

        def expm(self, X_s, dt):
            return self._blocks(lambda k: k.expm(dt)).view(Arr)

        def Q(self, dt, A_k, P_inf, X_spatial=None):
            return markov_Q(None, dt, A_k, P_inf)                                   # kernel.py:207-209
    return Kern, Prior


def _has_class(path, name):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name]
    return bool(cls) and any(isinstance(n, ast.FunctionDef) and n.name == "expm" for n in cls[0].body)


# -------------------------------------------------------------------------------------------- cases
sys.path.insert(0, OUT)
from cases import CASES  # noqa: E402


def synth(T, m, nan_frac, rng):
    t = onp.cumsum(rng.uniform(0.5, 1.5, T) * 0.1)
    Y = onp.sin(0.3 * onp.arange(T))[:, None] * rng.uniform(0.5, 1.5, m)[None] + 0.3 * rng.normal(size=(T, m))
    Y[rng.uniform(size=Y.shape) < nan_frac] = onp.nan
    G = rng.normal(size=(T, m, m)) * 0.2
    R = G @ onp.swapaxes(G, -1, -2) + 0.1 * onp.eye(m)
    return t, Y, R


def main():
    assert os.path.isdir(REF), "the reference tree is needed to (re)generate the golden vectors"
    jax = install_standin()
    sdes_mod = install_package_tree()
    settings = sys.modules["stgp.settings"]
    kf = importlib.import_module("stgp.computation.filters.kalman_filter")
    rts = importlib.import_module("stgp.computation.filters.rts_smoother")
    pkf = importlib.import_module("stgp.computation.filters.parallel_kalman_filter")
    prts = importlib.import_module("stgp.computation.filters.parallel_rts_smoother")
    eft = importlib.import_module("stgp.computation.natural_gradients.exponential_family_transforms")
    Kern, Prior = build_prior_classes(jax, sdes_mod)
    written = []

    for jitter in (1e-5, 0.0):
        settings.jitter = jitter
        for name, (latents, fso, T, nan_frac, seed) in CASES.items():
            rng = onp.random.default_rng(seed)
            prior = Prior([[Kern(*k) for k in lat] for lat in latents], fso)
            m = prior.H(None, None, None).shape[0]
            t, Y, R = synth(T, m, nan_frac, rng)
            data = types.SimpleNamespace(X_time=t.view(Arr), X_space=None, Nt=T, Ns=1, P=m,
                                         Y_st=Y[:, :, None].view(Arr))
            out = {"t": t, "Y": Y, "R": R, "jitter": jitter,
                   "P_inf": onp.asarray(prior.P_inf(None, None, None)), "H": onp.asarray(prior.H(None, None, None)),
                   "A_dt": onp.stack([onp.asarray(prior.expm(None, x)) for x in (0.0, 0.05, 0.9)])}
            # ---- reference sequential path: filter_loop -> evoke('filter','sequential'), smoother_loop
            lml, res = kf.filter_loop(data, prior, R=R.view(Arr), filter_type="sequential")
            out["seq_lml"], out["seq_mf"], out["seq_Pf"] = float(lml), onp.asarray(res["m"]), onp.asarray(res["P"])
            for fs in (False, True):
                mu, var = rts.smoother_loop(data, prior, res, full_state=fs, filter_type="sequential")
                out["seq_ms_full%d" % fs], out["seq_Ps_full%d" % fs] = onp.asarray(mu), onp.asarray(var)
            # ---- reference parallel path (bug-for-bug, whole-step masks only: SURVEY Q1/Q2)
            Yw = Y.copy()
            Yw[onp.isnan(Yw).any(axis=1)] = onp.nan           # whole-step missingness only
            data_w = types.SimpleNamespace(X_time=t.view(Arr), X_space=None, Nt=T, Ns=1, P=m,
                                           Y_st=Yw[:, :, None].view(Arr))
            try:
                lml_p, res_p = kf.filter_loop(data_w, prior, R=R.view(Arr), filter_type="parallel")
                mu_p, var_p = rts.smoother_loop(data_w, prior, res_p, full_state=False, filter_type="parallel")
                out["Y_wholestep"] = Yw
                out["par_lml"], out["par_mf"], out["par_Pf"] = float(lml_p), onp.asarray(res_p["m"]), onp.asarray(res_p["P"])
                out["par_ms"], out["par_Ps"] = onp.asarray(mu_p), onp.asarray(var_p)
            except Exception as e:                                                   # noqa: BLE001
                import traceback
                traceback.print_exc()
                print("  parallel path not runnable under the stand-in for %s: %r" % (name, e))
            fn = os.path.join(OUT, "filter_%s_jit%s.npz" % (name, "1e-5" if jitter else "0"))
            onp.savez_compressed(fn, **out)
            written.append(fn)

    # ---- CVI pieces: theta <-> lambda (real module), cvi_block_update and the block ELL (ast-extracted)
    settings.jitter = 1e-5
    ns = {"np": jax.numpy, "chex": sys.modules["chex"], "jit": _jit, "jax": jax, "settings": settings,
          "partial": __import__("functools").partial}
    mo = importlib.import_module("stgp.computation.matrix_ops")
    nu = importlib.import_module("stgp.utils.nan_utils")
    ga = importlib.import_module("stgp.computation.gaussian")
    for k in ("cholesky", "cholesky_solve", "add_jitter"):
        ns[k] = getattr(mo, k)
    for k in ("get_mask", "mask_to_identity", "mask_vector"):
        ns[k] = getattr(nu, k)
    for k in ("log_gaussian", "log_gaussian_with_nans", "log_gaussian_scalar"):
        if hasattr(ga, k):
            ns[k] = getattr(ga, k)
    blk = extract("computation/natural_gradients/cvi_nat_grad.py", ["cvi_block_update"], ns)["cvi_block_update"]
    ell = extract("computation/elbos/expected_log_likelihoods.py", ["full_gaussian_expected_log_likelihood"],
                  ns)["full_gaussian_expected_log_likelihood"]
    rng = onp.random.default_rng(11)
    cvi = {}
    for D in (1, 3, 6):
        for ngj in (1e-7, 1e-5):
            settings.ng_jitter = ngj
            G = rng.normal(size=(D, D))
            V = G @ G.T + 0.5 * onp.eye(D)
            Yt = rng.normal(size=(D, 1))
            l1, l2 = eft.theta_to_lambda(Yt.view(Arr), V.view(Arr))
            t1, t2 = eft.lambda_to_theta(l1, l2)
            G2 = rng.normal(size=(D, D))
            S = G2 @ G2.T + 0.2 * onp.eye(D)
            mq = rng.normal(size=(D, 1))
            dm = rng.normal(size=(D, 1))
            G3 = rng.normal(size=(D, D))
            dS = -(G3 @ G3.T) * 0.3
            beta = 0.37
            n1, n2 = blk(l1, l2, mq.view(Arr), S.view(Arr), dm.view(Arr), dS.view(Arr), beta, None)
            Yobs = rng.normal(size=(D, 1))
            if D > 1:
                Yobs[1, 0] = onp.nan
            e = ell(onp.zeros([D, 1]).view(Arr), Yobs.view(Arr), V.view(Arr), mq.view(Arr), S.view(Arr))
            key = "D%d_ngj%s" % (D, "1e-7" if ngj == 1e-7 else "1e-5")
            for nm, val in (("V", V), ("Yt", Yt), ("l1", l1), ("l2", l2), ("t1", t1), ("t2", t2), ("S", S), ("mq", mq),
                            ("dm", dm), ("dS", dS), ("beta", beta), ("n1", n1), ("n2", n2), ("Yobs", Yobs),
                            ("ell", e), ("ng_jitter", ngj)):
                cvi["%s_%s" % (key, nm)] = onp.asarray(val)
    fn = os.path.join(OUT, "cvi_blocks.npz")
    onp.savez_compressed(fn, **cvi)
    written.append(fn)
    for f in written:
        print("wrote", os.path.relpath(f, OUT), os.path.getsize(f), "bytes")


if __name__ == "__main__":
    main()
