"""CPU check of the CUDA kernels' register-level algebra: physs_core.cuh is compiled with g++ (the
same source the sm_100a kernels inline) and compared with the numpy oracle.  This is a test of the
product's source on the host, not a product code path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import filters, sde

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_core", "host_core.cpp")
SO = os.path.join(HERE, "host_core", "libhost_core.so")
CORE = os.path.join(os.path.dirname(HERE), "physs_gp_b200", "csrc", "physs_core.cuh")
P = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def lib():
    stale = (not os.path.exists(SO)
             or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)))
    if stale:
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                        "-o", SO, SRC], check=True)
    return ctypes.CDLL(SO)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(P)


def _blocks(kernels):
    out = []
    for k in kernels:
        out += k.parts if isinstance(k, sde.SumKernel) else [k]
    return out


CASES = [
    # kernels, given, hid, full_state_obs
    ([lambda: sde.Matern32(1.0, 1.3)], False, False, False),
    ([lambda: sde.Matern32(1.0, 1.3)], True, False, False),
    ([lambda: sde.Matern32(1.0, 1.3)], False, True, True),
    ([lambda: sde.Matern52(0.7, 1.3)], False, False, False),
    ([lambda: sde.Matern52(0.7, 1.3)], False, True, True),
    ([lambda: sde.Matern72(1.2, 0.9)], False, False, False),
    ([lambda: sde.Matern72(1.2, 0.9)], True, False, True),
    ([lambda: sde.Matern72(1.2, 0.9)], False, True, True),
    ([lambda: sde.SumKernel([sde.Matern32(1.0, 1.3), sde.Matern32(0.4, 0.5)])], False, False, False),
    ([lambda: sde.Matern32(1.0, 1.3), lambda: sde.Matern32(0.4, 0.5)], False, False, False),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("jitter", [1e-5, 0.0])
def test_core_matches_oracle(lib, case, jitter):
    mk, given, hid, fso = CASES[case]
    kernels = [f() for f in mk]
    rng = np.random.default_rng(case)
    prior = (sde.LTI_SDE_Full_State_Obs if fso else sde.LTI_SDE)(kernels)
    d = prior.state_dim
    H = np.ascontiguousarray(prior.H())
    m = H.shape[0]
    T = 200
    t = np.cumsum(rng.uniform(0.5, 1.5, T)) * 0.1
    Y = rng.normal(size=(T, m))
    Y[rng.uniform(size=(T, m)) < 0.1] = np.nan
    G = rng.normal(size=(T, m, m)) * 0.2
    R = G @ G.transpose(0, 2, 1) + 0.1 * np.eye(m)
    lml, mf, Pf, lk = filters.filter_sequential(prior, t, Y, R, jitter)
    ms, Ps = filters.smoother_sequential(prior, t, mf, Pf, full_state=True, jitter=jitter)
    dt = np.hstack([0, np.diff(t)])
    dts = np.hstack([np.diff(t), 0])
    Pinf = prior.P_inf()
    m0 = prior.m_inf()[:, 0].copy()
    blocks = _blocks(kernels)
    s = blocks[0].state_dim
    lam = np.array([np.sqrt(2.0 * b.state_dim - 1.0) / b.ls for b in blocks])
    A = Q = As = Qs = None
    if given:
        A = np.array([prior.expm(x) for x in dt])
        Q = np.array([Pinf - a @ Pinf @ a.T for a in A])
        As = np.array([prior.expm(x) for x in dts])
        Qs = np.array([Pinf - a @ Pinf @ a.T for a in As])
        s = d
    mf2, Pf2, lk2 = np.zeros((T, d)), np.zeros((T, d, d)), np.zeros(T)
    ltot = np.zeros(1)
    rc = lib.host_filter(d, s, m, int(hid), int(given), ctypes.c_int64(T), _ptr(A), _ptr(Q), _ptr(lam),
                         _ptr(dt), _ptr(Pinf), _ptr(m0), _ptr(Pinf), _ptr(H), _ptr(Y), _ptr(R),
                         ctypes.c_int64(m * m), ctypes.c_double(jitter), _ptr(mf2), _ptr(Pf2), _ptr(lk2),
                         _ptr(ltot))
    assert rc == 0
    ms2, Ps2 = np.zeros((T, d)), np.zeros((T, d, d))
    rc = lib.host_smooth(d, s, int(given), ctypes.c_int64(T), _ptr(As), _ptr(Qs), _ptr(lam), _ptr(dts),
                         _ptr(Pinf), _ptr(mf2), _ptr(Pf2), ctypes.c_double(jitter), _ptr(ms2), _ptr(Ps2))
    assert rc == 0

    def rel(a, b):
        return np.abs(a - b).max() / np.abs(b).max()

    tol = 1e-11
    assert rel(mf2, mf[:, :, 0]) < tol
    assert rel(Pf2, Pf) < tol
    assert abs(lk2.sum() - lml) < tol * abs(lml)
    assert abs(ltot[0] - lml) < tol * abs(lml)      # product-of-determinants accumulator
    assert rel(lk2, lk) < 1e-10
    assert rel(ms2, ms[:, :, 0]) < tol
    assert rel(Ps2, Ps) < tol
