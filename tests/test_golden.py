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
