"""ctypes wrapper of oracle/ssm_oracle.c (TEST INFRASTRUCTURE, see oracle/__init__.py).

`build()` compiles it with gcc -O2 -fopenmp into oracle/libssm_oracle.so (git-ignored, travels to the
GPU box with the snapshot).  -ffp-contract=off keeps gcc from fusing multiply-adds so the C oracle
rounds like the numpy one."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ssm_oracle.c")
SO = os.path.join(HERE, "libssm_oracle.so")
_P = ctypes.POINTER(ctypes.c_double)
_lib = None


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC",
                        "-o", SO, SRC, "-lm"], check=True)
    return SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = ctypes.CDLL(SO)
        _lib.oracle_filter_smooth.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


def filter_smooth(block_size, lam, Pinf, H, t, Y, R, jitter=1e-5, full_state=True, smooth=True,
                  keep_filtered=True, nthreads=0):
    """lam [B|1, nblk]; Pinf [B|1, d, d]; H [m, d]; t [T]; Y [B, T, m]; R [m, m] | [T, m, m] | [B, T, m, m].
    Returns dict(lml [B], mf, Pf, ms, Ps, threads)."""
    lib = _load()
    lam = np.ascontiguousarray(np.atleast_2d(lam), dtype=np.float64)
    Pinf = np.ascontiguousarray(Pinf, dtype=np.float64)
    if Pinf.ndim == 2:
        Pinf = Pinf[None]
    H = np.ascontiguousarray(H, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B, T, m = Y.shape
    d = Pinf.shape[-1]
    nblk = lam.shape[1]
    assert nblk * block_size == d and H.shape == (m, d)
    R = np.ascontiguousarray(R, dtype=np.float64)
    if R.ndim == 2:
        R_bs, R_ts = 0, 0
    elif R.ndim == 3:
        R_bs, R_ts = 0, m * m
    else:
        R_bs, R_ts = T * m * m, m * m
    dt_f = np.ascontiguousarray(np.hstack([0.0, np.diff(t)]))
    dt_s = np.ascontiguousarray(np.hstack([np.diff(t), 0.0]))
    mp = d if full_state else m
    mf = np.empty((B, T, d)) if keep_filtered else None
    Pf = np.empty((B, T, d, d)) if keep_filtered else None
    ms = np.empty((B, T, mp)) if smooth else None
    Ps = np.empty((B, T, mp, mp)) if smooth else None
    lml = np.empty(B)
    i64 = ctypes.c_int64
    used = lib.oracle_filter_smooth(
        i64(B), i64(T), d, m, block_size, nblk, _p(lam), i64(nblk if lam.shape[0] > 1 else 0),
        _p(Pinf), i64(d * d if Pinf.shape[0] > 1 else 0), _p(H), _p(dt_f), _p(dt_s), _p(Y), _p(R),
        i64(R_bs), i64(R_ts), ctypes.c_double(jitter), int(full_state), _p(mf), _p(Pf), _p(lml),
        _p(ms), _p(Ps), int(nthreads))
    assert used > 0
    return dict(lml=lml, mf=mf, Pf=Pf, ms=ms, Ps=Ps, threads=used)
