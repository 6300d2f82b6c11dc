"""Pins the oracle (numpy restatement and C port) against tests/golden/: outputs of the reference's OWN
source files, produced by tests/golden/make_golden.py (see tests/golden/README.md).  Runs on CPU.

Tolerance: 1e-12 relative (array scale) for everything evaluated in the same operation order; 1e-9 for
the parallel-scan path, where the oracle folds left-to-right and the golden vectors follow
jax.lax.associative_scan's odd/even tree (with jitter != 0 the jittered gains make the operator only
approximately associative, so re-association moves the result at the 1e-10 level)."""
import glob
import os

import numpy as np
import pytest

from oracle import cvi as ocvi
from oracle import filters as ofilters
from oracle import sde as osde
from tests.golden.cases import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KIND = {"m32": osde.Matern32, "m52": osde.Matern52, "m72": osde.Matern72}


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def oracle_prior(name):
    latents, fso = CASES[name][0], CASES[name][1]
    lat = []
    for parts in latents:
        ks = [KIND[k](ls, var) for k, ls, var in parts]
        lat.append(ks[0] if len(ks) == 1 else osde.SumKernel(ks))
    return osde.LTI_SDE_Full_State_Obs(lat) if fso else osde.LTI_SDE(lat)


def load(name, jit):
    return np.load(os.path.join(GOLD, "filter_%s_jit%s.npz" % (name, "1e-5" if jit else "0")))


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "filter_*.npz"))) == 2 * len(CASES)
    assert os.path.exists(os.path.join(GOLD, "cvi_blocks.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_discretisation_matches_reference(name):
    g = load(name, 1e-5)
    prior = oracle_prior(name)
    assert rel(prior.P_inf(), g["P_inf"]) < 1e-14
    assert rel(prior.H(), g["H"]) == 0.0
    for i, dt in enumerate((0.0, 0.05, 0.9)):
        assert rel(prior.expm(dt), g["A_dt"][i]) < 1e-13


@pytest.mark.parametrize("jit", [1e-5, 0.0])
@pytest.mark.parametrize("name", sorted(CASES))
def test_sequential_filter_smoother_matches_reference(name, jit):
    g = load(name, jit)
    prior = oracle_prior(name)
    lml, mf, Pf, _ = ofilters.filter_sequential(prior, g["t"], g["Y"], g["R"], jit)
    assert abs(lml - float(g["seq_lml"])) <= 1e-12 * abs(float(g["seq_lml"]))
    assert rel(mf, g["seq_mf"]) < 1e-12 and rel(Pf, g["seq_Pf"]) < 1e-12
    for fs in (False, True):
        ms, Ps = ofilters.smoother_sequential(prior, g["t"], mf, Pf, full_state=fs, jitter=jit)
        assert rel(ms, g["seq_ms_full%d" % fs]) < 1e-12
        assert rel(Ps, g["seq_Ps_full%d" % fs]) < 1e-12


@pytest.mark.parametrize("jit", [1e-5, 0.0])
@pytest.mark.parametrize("name", sorted(CASES))
def test_parallel_reference_path_matches_reference(name, jit):
    """filter('parallel') / smoother('parallel') bug-for-bug (quirks Q1-Q3), whole-step masks."""
    g = load(name, jit)
    prior = oracle_prior(name)
    lml, mf, Pf = ofilters.filter_parallel_reference(prior, g["t"], g["Y_wholestep"], g["R"], jit)
    assert abs(lml - float(g["par_lml"])) <= 1e-9 * abs(float(g["par_lml"]))
    assert rel(mf, g["par_mf"]) < 1e-9 and rel(Pf, g["par_Pf"]) < 1e-9
    ms, Ps = ofilters.smoother_parallel_reference(prior, g["t"], g["par_mf"], g["par_Pf"], jit)
    assert rel(ms, g["par_ms"]) < 1e-9 and rel(Ps, g["par_Ps"]) < 1e-9


@pytest.mark.parametrize("name", ["m32", "m52", "m72", "sum_m32_m52"])
def test_c_port_matches_reference(name):
    """oracle/ssm_oracle.c (the CPU baseline of bench.py) on the golden inputs (time-invariant R only)."""
    from oracle import c_oracle
    c_oracle.build()
    g = load(name, 1e-5)
    prior = oracle_prior(name)
    latents = CASES[name][0]
    sizes = {osde.Matern32: 2, osde.Matern52: 3, osde.Matern72: 4}
    kinds = [KIND[k] for parts in latents for k, _, _ in parts]
    if len({sizes[k] for k in kinds}) != 1:
        pytest.skip("the C port takes equal-size Matern blocks")
    s = sizes[kinds[0]]
    lam = np.array([[np.sqrt(2 * (s - 0.5)) / ls for parts in latents for _, ls, _ in parts]])
    R0 = g["R"][0]
    lml_o, mf_o, Pf_o, _ = ofilters.filter_sequential(prior, g["t"], g["Y"], np.tile(R0, [len(g["t"]), 1, 1]), 1e-5)
    out = c_oracle.filter_smooth(s, lam, prior.P_inf()[None], prior.H(), g["t"], g["Y"][None], R0,
                                 jitter=1e-5, full_state=True, keep_filtered=True)
    assert rel(out["mf"][0], mf_o[..., 0]) < 1e-12 and rel(out["Pf"][0], Pf_o) < 1e-12
    assert abs(out["lml"][0] - lml_o) <= 1e-12 * abs(lml_o)


def test_cvi_blocks_match_reference():
    g = np.load(os.path.join(GOLD, "cvi_blocks.npz"))
    for D in (1, 3, 6):
        for tag, ngj in (("1e-7", 1e-7), ("1e-5", 1e-5)):
            k = "D%d_ngj%s_" % (D, tag)
            l1, l2 = ocvi.theta_to_lambda(g[k + "Yt"], g[k + "V"], ngj)
            assert rel(l1, g[k + "l1"]) < 1e-12 and rel(l2, g[k + "l2"]) < 1e-12
            t1, t2 = ocvi.lambda_to_theta(g[k + "l1"], g[k + "l2"], ngj)
            assert rel(t1, g[k + "t1"]) < 1e-12 and rel(t2, g[k + "t2"]) < 1e-12
            n1, n2 = ocvi.cvi_block_update(g[k + "l1"], g[k + "l2"], g[k + "mq"], g[k + "S"], g[k + "dm"],
                                           g[k + "dS"], float(g[k + "beta"]))
            assert rel(n1, g[k + "n1"]) < 1e-13 and rel(n2, g[k + "n2"]) < 1e-13
            e = ocvi.full_gaussian_ell(g[k + "Yobs"], g[k + "V"], g[k + "mq"], g[k + "S"])
            assert abs(e - float(g[k + "ell"])) <= 1e-12 * abs(float(g[k + "ell"]))


def test_cvi_precision_blocks_match_reference():
    """'NG_Precision' (tests/golden/make_golden_prec.py): the reference's own theta_precision_to_lambda ->
    cvi_block_update -> lambda_to_theta_precision, and mat_inv of the stored precision."""
    from oracle import linalg as ola
    g = np.load(os.path.join(GOLD, "cvi_blocks_prec.npz"))
    for D in (1, 2, 3, 4, 6, 8):
        for tag, ngj in (("1e-7", 1e-7), ("1e-5", 1e-5)):
            k = "D%d_ngj%s_" % (D, tag)
            l1, l2 = ocvi.theta_precision_to_lambda(g[k + "Yt"], g[k + "Lam"], ngj)
            assert rel(l1, g[k + "l1"]) < 1e-12 and rel(l2, g[k + "l2"]) < 1e-12
            t1, t2 = ocvi.lambda_to_theta_precision(g[k + "n1"], g[k + "n2"], ngj)
            assert rel(t1, g[k + "t1"]) < 1e-12 and rel(t2, g[k + "t2"]) < 1e-12
            Yn, Pn = ocvi.cvi_step_precision(g[k + "Yt"][None, :, 0], g[k + "Lam"][None], g[k + "mq"][None, :, 0],
                                             g[k + "S"][None], g[k + "dm"][None, :, 0], g[k + "dS"][None],
                                             float(g[k + "beta"]), ngj)
            assert rel(Yn[0], g[k + "t1"][:, 0]) < 1e-12 and rel(Pn[0], g[k + "t2"]) < 1e-12
            assert rel(ola.mat_inv(g[k + "Lam"], 1e-5), g[k + "var"]) < 1e-12


# ---------------------------------------------------------------------------------------------------------------
# tests/golden/make_golden_cvi.py: the reference's OWN natural_gradients / elbo / Gauss-Newton assembly /
# Independent stacking, executed in place.  These move rows a10 / a12 / a13 / a6 from "oracle compared with
# nothing" to "oracle pinned to reference output".
def _assembly_cases():
    g = np.load(os.path.join(GOLD, "cvi_assembly.npz"))
    return g, sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_elbo")})


def _assembly_prior(g, key):
    kind = str(g[key + "_kernel"][0])
    ls, var = g[key + "_hyper"]
    return osde.LTI_SDE_Full_State_Obs([KIND[kind](float(ls), float(var))])


def test_oracle_natural_gradients_and_elbo_match_reference_assembly():
    """oracle/cvi.py composed as in cvi_nat_grad.py:346-410 + cvi_parameterisations.py:63-93 + elbos.py:163-194
    == the reference's own `natural_gradients` and `elbo` (Gaussian likelihood, full-state sites, missing data)."""
    g, keys = _assembly_cases()
    assert len(keys) == 4
    for key in keys:
        prior = _assembly_prior(g, key)
        t, Ytil, Vtil, Yobs, noise = (g[key + "_" + n] for n in ("t", "Ytil", "Vtil", "Yobs", "noise"))
        beta, ngj, jit = float(g[key + "_beta"]), float(g[key + "_ng_jitter"]), float(g[key + "_jitter"])
        T, D = Ytil.shape

        def posterior(Yt, Vt):
            lml, mf, Pf, _ = ofilters.filter_sequential(prior, t, Yt, Vt, jit)
            ms, Ps = ofilters.smoother_sequential(prior, t, mf, Pf, full_state=False, jitter=jit)
            return lml, ms[..., 0], Ps
        lml, q_mu, q_var = posterior(Ytil, Vtil)
        assert rel(q_mu, g[key + "_q_mu"][..., 0]) < 1e-11 and rel(q_var, g[key + "_q_var"][:, 0]) < 1e-11
        grads = [ocvi.gaussian_ell_and_grads(Yobs[k], noise, None, q_mu[k], q_var[k]) for k in range(T)]
        dm = np.stack([x[1] for x in grads])
        dS = np.stack([x[2] for x in grads])
        Yn, Vn = ocvi.cvi_step(Ytil, Vtil, q_mu, q_var, dm, dS, beta, ng_jitter=ngj)
        # the reference's gradients come from central differences in the generator (exact for the quadratic
        # Gaussian ELL up to ~1e-12 round-off), the site inversion amplifies by cond(V~): 1e-9 stated
        assert rel(Yn, g[key + "_Ytil_new"]) < 1e-9, key
        assert rel(Vn, g[key + "_Vtil_new"]) < 1e-9, key
        ell = sum(x[0] for x in grads)
        ell_s = ocvi.surrogate_ell(Ytil, Vtil, q_mu, q_var)
        assert abs(ocvi.elbo(ell, ell_s, lml) - float(g[key + "_elbo"])) < 1e-10 * abs(float(g[key + "_elbo"])), key


def test_oracle_gauss_newton_curvature_matches_reference_assembly():
    """oracle pendulum Gauss-Newton curvature (delta-u, delta-f, Laplace) == 0.5 * the reference's own
    mask / J^T (-1/var) J / sum lines (cvi_hessian_approximations.py, slice recorded in the fixture) on the
    Jacobian of the reference's DampedPendulum1D.forward (complex-step)."""
    g = np.load(os.path.join(GOLD, "gn_pendulum.npz"))
    u, Y, H = g["u"], g["Y"], g["approx_hessian"]
    gl, b = float(g["g"]) / float(g["l"]), float(g["b"])
    for k in range(u.shape[0]):
        r = ocvi.pendulum_forward(u[k], gl, b)
        assert abs(r[1] - g["residual"][k]) < 1e-13 * max(1.0, abs(g["residual"][k]))
        _, _, dS = ocvi.pendulum_ell_and_grads(Y[k], u[k], np.eye(4) * 0.1, gl, b, float(g["var_obs"]),
                                               float(g["var_col"]), gauss_newton=True)
        assert rel(dS, H[k]) < 1e-13 or np.abs(dS - H[k]).max() < 1e-13


def test_oracle_independent_stacking_matches_reference():
    """oracle/sde.py LTI_SDE stacking == the reference's Independent.{expm, P_inf, H, m_inf, Q}
    (transforms/transform.py:400-545 with matrix_ops.to_block_diag / get_block_diagonal)."""
    g = np.load(os.path.join(GOLD, "independent_stack.npz"))
    for name in ("m32_m32", "m52x3"):
        kinds, hyper = g[name + "_kinds"], g[name + "_hyper"]
        prior = osde.LTI_SDE([KIND[str(k)](float(h[0]), float(h[1])) for k, h in zip(kinds, hyper)])
        assert rel(prior.H(), g[name + "_H"]) == 0.0
        assert np.abs(prior.m_inf()[:, 0] - np.ravel(g[name + "_minf"])).max() == 0.0
        assert rel(prior.P_inf(), g[name + "_Pinf_ssr"]) < 1e-14
        for dt in (0.05, 0.9):
            key = "%s_dt%s" % (name, dt)
            Ak = prior.expm(dt)
            assert rel(Ak, g[key + "_A"]) < 1e-14 and rel(prior.P_inf(), g[key + "_Pinf"]) < 1e-14
            assert rel(prior.Q(dt, Ak, prior.P_inf()), g[key + "_Q"]) < 1e-12


def test_product_prior_stacking_matches_reference():
    """The PRODUCT's host mirror (physs_gp_b200.sdes / kernels, numpy only) against the same reference vectors."""
    from physs_gp_b200 import kernels as K
    from physs_gp_b200 import sdes
    g = np.load(os.path.join(GOLD, "independent_stack.npz"))
    PK = {"m32": K.Matern32, "m52": K.Matern52, "m72": K.Matern72}
    for name in ("m32_m32", "m52x3"):
        kinds, hyper = g[name + "_kinds"], g[name + "_hyper"]
        prior = sdes.LTI_SDE(sdes.Independent([PK[str(k)](float(h[0]), float(h[1])) for k, h in zip(kinds, hyper)]))
        assert rel(prior.H(), g[name + "_H"]) == 0.0
        for dt in (0.05, 0.9):
            key = "%s_dt%s" % (name, dt)
            Ak = prior.expm(None, dt)
            Pinf = prior.P_inf()
            assert rel(Ak, g[key + "_A"]) < 1e-14 and rel(Pinf, g[key + "_Pinf"]) < 1e-14
            assert rel(prior.Q(dt, Ak, Pinf), g[key + "_Q"]) < 1e-12


# ------------------------------------------------------------------------------- collocation (EKF) filter step
def _ekf_files():
    return sorted(glob.glob(os.path.join(GOLD, "ekf_*.npz")) + glob.glob(os.path.join(GOLD, "ekfsys_*.npz")))


def ekf_residuals(g):
    out = []
    for p in range(int(g["n_res"])):
        terms = [(str(k), int(i), float(c)) for k, i, c in
                 zip(g["term_kind%d" % p], g["term_idx%d" % p], g["term_coef%d" % p])]
        f = g["forcing%d" % p]
        out.append(ofilters.PointResidual(g["w%d" % p], terms, f if f.size else None))
    return out


def test_ekf_golden_files_present():
    assert len(_ekf_files()) == 7


@pytest.mark.parametrize("path", _ekf_files(), ids=[os.path.basename(p)[:-4] for p in _ekf_files()])
def test_oracle_collocation_filter_matches_reference(path):
    """oracle.filters.filter_pde_sequential == the reference's kf_predict_step(PDE, 'sequential')
    (kalman_filter.py:340-427) run by tests/golden/make_golden_ekf.py, and the smoother on its output ==
    rts_step_wrapper(PDE) (rts_smoother.py:108-150)."""
    g = np.load(path)
    hyper = np.atleast_2d(g["hyper"])               # one row per latent (ekfsys_*: systems of ODEs over 2 / 3 latents)
    prior = osde.LTI_SDE([KIND[str(k)](float(ls), float(var)) for k, (ls, var) in zip(g["kernel"], hyper)])
    jit = float(g["jitter"])
    bnd = g["boundary"] if "boundary" in g.files else None
    lml, mf, Pf, _ = ofilters.filter_pde_sequential(prior, ekf_residuals(g), g["t"], g["Y"], g["R"], boundary=bnd,
                                                    y_pseudo=g["y_pseudo"], observe_data=bool(g["observe_data"]),
                                                    jitter=jit)
    assert abs(lml - float(g["lml"])) < 1e-11 * abs(float(g["lml"]))
    assert rel(mf, g["mf"]) < 1e-11 and rel(Pf, g["Pf"]) < 1e-11
    ms, Ps = ofilters.smoother_sequential(prior, g["t"], mf, Pf, full_state=True, jitter=jit)
    assert rel(ms, g["ms"]) < 1e-10 and rel(Ps, g["Ps"]) < 1e-10


def _check_sort_pad(device):
    import torch
    from physs_gp_b200.data import SequentialData
    g = np.load(os.path.join(GOLD, "sort_pad.npz"))
    X = torch.as_tensor(g["X"], device=device)
    Y = torch.as_tensor(g["Y"], device=device)
    sd = SequentialData()
    Xs, Yst = sd.sort(X, Y)
    assert sd.num_points_added == int(g["points_added"])
    assert np.array_equal(Xs.cpu().numpy(), g["X_sorted"])
    assert np.array_equal(np.nan_to_num(Yst.cpu().numpy(), nan=-777.0), np.nan_to_num(g["Y_st"], nan=-777.0))
    assert np.array_equal(sd.unique_idx.cpu().numpy(), g["unique_idx"])
    assert np.array_equal(sd.reverse_unique_idx.cpu().numpy(), g["reverse_idx"])
    out = sd.unsort(torch.as_tensor(g["payload"], device=device))
    assert np.array_equal(out.cpu().numpy(), g["unsorted"])


def test_device_sort_pad_matches_reference_on_cpu_tensors():
    """physs_gp_b200.data.SequentialData (torch) == the reference's pad_with_nan_to_make_grid +
    order_sequentially_np + unsort (data/sequential.py:9-144, data/data.py:353-415; fixture from
    tests/golden/make_golden_cvi.py): scattered, shuffled, duplicated space-time points."""
    _check_sort_pad("cpu")


# --------------------------------------------------------------------------- integrated Wiener prior (row a6)
@pytest.mark.parametrize("q", [1, 2, 3])
def test_oracle_and_product_iwp_match_reference(q):
    """oracle.sde.IWP and the product's kernels.WienerVelocity == the reference's WienerVelocity.{expm, Q, to_ss}
    (kernels/wiener.py:90-149), and the oracle filter / smoother on that prior == the reference's
    (tests/golden/make_golden_ekf.py -> iwp_q*.npz)."""
    from physs_gp_b200 import kernels as K
    g = np.load(os.path.join(GOLD, "iwp_q%d.npz" % q))
    var, ssc, jit = float(g["variance"]), float(g["stable_state_covariance"]), float(g["jitter"])
    ok = osde.IWP(q, var, ssc)
    pk = K.WienerVelocity(q, var, ssc)
    for i, dt in enumerate((0.0, 0.05, 0.9)):
        for k in (ok, pk):
            assert rel(k.expm(dt), g["A_dt"][i]) < 1e-14
            Q = k.Q(dt)
            assert np.abs(Q - g["Q_dt"][i]).max() <= 1e-14 * max(np.abs(g["Q_dt"][i]).max(), 1e-300)
    prior = osde.LTI_SDE([ok])
    lml, mf, Pf, _ = ofilters.filter_sequential(prior, g["t"], g["Y"], g["R"], jit)
    assert abs(lml - float(g["lml"])) < 1e-11 * abs(float(g["lml"]))
    assert rel(mf, g["mf"]) < 1e-11 and rel(Pf, g["Pf"]) < 1e-11
    for fs in (False, True):
        ms, Ps = ofilters.smoother_sequential(prior, g["t"], mf, Pf, full_state=fs, jitter=jit)
        assert rel(ms, g["ms_full%d" % fs]) < 1e-10 and rel(Ps, g["Ps_full%d" % fs]) < 1e-10


# --------------------------------------------------------------------------- periodic prior (row a6)
def _periodic_files():
    return sorted(glob.glob(os.path.join(GOLD, "periodic_*.npz")))


def periodic_priors(g, osde_mod=None, K=None):
    """(oracle kernel list, product kernel) of one periodic golden case."""
    args = (float(g["frequency"]), float(g["lengthscale"]), float(g["variance"]), int(g["n_terms"]))
    extra = g["extra_m32"]
    ok = pk = None
    if osde_mod is not None:
        ok = [osde_mod.ApproxPeriodicBN(*args)] + ([osde_mod.Matern32(*extra)] if extra.size else [])
        ok = [osde_mod.SumKernel(ok)] if len(ok) > 1 else ok
    if K is not None:
        pk = K.ApproxSDEPeriodic_BN(*args)
        if extra.size:
            pk = K.SumKernel(pk, K.Matern32(extra[0], extra[1]))
    return ok, pk


def test_periodic_golden_files_present():
    assert len(_periodic_files()) == 5


@pytest.mark.parametrize("path", _periodic_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_and_product_periodic_match_reference(path):
    """oracle.sde.ApproxPeriodicBN (generic Pade expm, as the reference) and the product's closed-form rotation stack
    kernels.ApproxSDEPeriodic_BN == the reference's ApproxSDEPeriodic_BN.{to_ss, expm} (kernels/periodic.py:213-253),
    and the oracle filter / smoother on that prior == the reference's (tests/golden/make_golden_periodic.py)."""
    from physs_gp_b200 import kernels as K, sdes
    g = np.load(path)
    ok, pk = periodic_priors(g, osde, K)
    prior = osde.LTI_SDE(ok)
    pprior = sdes.LTI_SDE(sdes.Independent([pk]))
    assert rel(prior.P_inf(), g["P_inf"]) < 1e-14 and rel(prior.H(), g["H"]) == 0.0
    assert rel(pprior.P_inf(), g["P_inf"]) < 1e-14 and rel(pprior.H(), g["H"]) == 0.0
    for i, dt in enumerate((0.0, 0.05, 0.9)):
        assert rel(prior.expm(dt), g["A_dt"][i]) < 1e-13
        # closed-form rotations vs the reference's Pade expm (whose own error reaches 2e-13 at j w dt = 56 rad)
        assert rel(pprior.expm(None, dt), g["A_dt"][i]) < 1e-12
    blocks = pprior.ss_blocks()
    assert all(s == 2 for s, _ in blocks) and np.signbit(blocks[0][1]) and blocks[0][1] == 0.0
    jit = float(g["jitter"])
    lml, mf, Pf, _ = ofilters.filter_sequential(prior, g["t"], g["Y"], g["R"], jit)
    assert abs(lml - float(g["lml"])) <= 1e-10 * abs(float(g["lml"]))
    assert rel(mf, g["mf"]) < 1e-9 and rel(Pf, g["Pf"]) < 1e-9
    for fs in (False, True):
        ms, Ps = ofilters.smoother_sequential(prior, g["t"], mf, Pf, full_state=fs, jitter=jit)
        assert rel(ms, g["ms_full%d" % fs]) < 1e-8 and rel(Ps, g["Ps_full%d" % fs]) < 1e-8


# ---- separable spatio-temporal prior (config 2): tests/golden/make_golden_st.py
def _st_files():
    return sorted(glob.glob(os.path.join(GOLD, "st_*.npz")))


def test_st_golden_files_present():
    assert len(_st_files()) == 4


@pytest.mark.parametrize("path", _st_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_separable_prior_matches_reference(path):
    """oracle/sde.py:SpaceTimeSeparable (Kronecker stacking of kernel.py:213-265 / ss_utils.py:41-53) and the oracle's
    filter / smoother on it against what the reference's own classes produce."""
    g = np.load(path)
    kind = osde.Matern32 if "m32" in path else osde.Matern52
    prior = osde.LTI_SDE([osde.SpaceTimeSeparable(kind(*g["temporal"]), g["Ks"])])
    assert rel(prior.P_inf(), g["P_inf"]) < 1e-14 and rel(prior.H(), g["H"]) == 0.0
    for i, dt in enumerate((0.0, 0.05, 0.9)):
        assert rel(prior.expm(dt), g["A_dt"][i]) < 1e-13
    jit = float(g["jitter"])
    lml, mf, Pf, _ = ofilters.filter_sequential(prior, g["t"], g["Y"], g["R"], jit)
    assert abs(lml - float(g["seq_lml"])) <= 1e-11 * abs(float(g["seq_lml"]))
    assert rel(mf, g["seq_mf"]) < 1e-10 and rel(Pf, g["seq_Pf"]) < 1e-10
    for fs in (False, True):
        ms, Ps = ofilters.smoother_sequential(prior, g["t"], mf, Pf, full_state=fs, jitter=jit)
        assert rel(ms, g["seq_ms_full%d" % fs]) < 1e-9 and rel(Ps, g["seq_Ps_full%d" % fs]) < 1e-9


# ---- spatial conditional after the smoother (row f3): tests/golden/make_golden_spatial.py
def _spatial_files():
    return sorted(glob.glob(os.path.join(GOLD, "spatial_cond_*.npz")))


def test_spatial_golden_files_present():
    assert len(_spatial_files()) == 3


@pytest.mark.parametrize("path", _spatial_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_spatial_conditional_matches_reference(path):
    """oracle/dense_gp.py:spatial_conditional and the product's host-side weights (W, C0) against the reference's own
    gaussian_spatial_conditional_cholesky vmapped over time (marginals.py:82-113)."""
    from oracle import dense_gp
    from physs_gp_b200 import spatial
    g = np.load(path)
    jit = float(g["jitter"])
    mu, var = dense_gp.spatial_conditional(g["Kzz"], g["Ksz"], g["Kss"], g["Ktt"], g["pred_mean"], g["pred_var"], jit)
    assert rel(mu, g["mu"]) < 1e-12 and rel(var, g["var"]) < 1e-12
    # the closed form the kernel evaluates: var_t = ktt_t C0 + W (P_t + jitter I) W^T
    W, C0 = spatial.conditional_weights(g["Kzz"], g["Ksz"], g["Kss"], jit)
    M = W.shape[1]
    for t in range(g["pred_mean"].shape[0]):
        v = g["Ktt"][t] * C0 + W @ (g["pred_var"][t] + jit * np.eye(M)) @ W.T
        assert rel(v, g["var"][t, 0]) < 1e-11
        assert rel(W @ g["pred_mean"][t], g["mu"][t]) < 1e-12
