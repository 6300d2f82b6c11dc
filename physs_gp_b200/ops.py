"""Device-level entry points: thin, allocation-explicit wrappers over the C ABI
(include/physs_b200.h) on torch CUDA tensors.  torch is plumbing here (device memory, streams);
every FLOP of the path runs in libphyss_b200.so.  No CPU fallback: non-CUDA tensors raise.

Broadcasting: every argument is viewed with `expand` to its full [B, ...] shape and the resulting
batch (and, for R, time) stride is handed to the kernel; a stride of 0 means "shared".
"""
import os

import numpy as np
import torch

from . import _lib
from . import settings


class Disc:
    """How A_k = expm(F dt_k), Q_k reach the kernels.

    Disc.given(A, Q):            A, Q broadcastable to [B, T, d, d]  (PHYSS_DISC_GIVEN)
    Disc.matern(nblk, lam, Pinf): nblk equal-size Matern blocks, lam -> [B, nblk], Pinf -> [B, d, d]
                                 block-diagonal stationary covariance (PHYSS_DISC_MATERN)
    """

    def __init__(self, mode, nblk=0, A=None, Q=None, lam=None, Pinf=None):
        self.mode, self.nblk, self.A, self.Q, self.lam, self.Pinf = mode, nblk, A, Q, lam, Pinf

    @staticmethod
    def given(A, Q):
        return Disc(_lib.DISC_GIVEN, 0, A=A, Q=Q)

    @staticmethod
    def matern(nblk, lam, Pinf):
        return Disc(_lib.DISC_MATERN, int(nblk), lam=lam, Pinf=Pinf)

    @staticmethod
    def iwp(var):
        """One integrated-Wiener block (PHYSS_DISC_IWP): var [B, 1] = spectral density per series."""
        return Disc(_lib.DISC_IWP, 1, lam=var)


def _dev(x, name):
    if not isinstance(x, torch.Tensor):
        raise TypeError("%s must be a torch tensor" % name)
    if not x.is_cuda:
        raise RuntimeError("%s is not on a CUDA device; physs_gp_b200 has no CPU path" % name)
    if x.dtype != torch.float64:
        raise TypeError("%s must be float64" % name)
    return x


def _bview(x, name, shape, inner):
    """Expand x to `shape`; require the trailing `inner` dims dense; return (tensor, strides)."""
    x = _dev(x, name)
    try:
        v = x.expand(*shape)
    except RuntimeError:
        raise ValueError("%s with shape %s is not broadcastable to %s" % (name, tuple(x.shape), tuple(shape)))
    exp = 1
    ok = True
    for dim in range(len(shape) - 1, len(shape) - 1 - inner, -1):
        if shape[dim] != 1 and v.stride(dim) != exp:
            ok = False
        exp *= shape[dim]
    if not ok:
        # materialise the trailing `inner` dims (the kernels index them densely: a broadcast A_k [d, d] or
        # dt [1] must become real [T, d, d] / [T] storage); the leading dims may stay broadcast (stride 0)
        lead = len(shape) - inner
        xs = x.reshape((1,) * (len(shape) - x.dim()) + tuple(x.shape))
        xs = xs.expand(*(tuple(xs.shape[:lead]) + tuple(shape[lead:]))).contiguous()
        v = xs.expand(*shape)
    return v, v.stride()


def step_layout(x, name):
    """(tensor, time_major) for a per-step array x [B, T, ...]: the kernels take it batch-major (contiguous
    [B, T, ...]) or time-major (a transposed view of a contiguous [T, B, ...] tensor) without a copy;
    anything else is made batch-major contiguous."""
    x = _dev(x, name)
    if x.is_contiguous():
        return x, False
    if x.transpose(0, 1).is_contiguous():
        return x, True
    return x.contiguous(), False


def empty_steps(B, T, inner, dev, time_major):
    """Uninitialised per-step array of logical shape [B, T, *inner] in the requested memory order."""
    if time_major:
        return torch.empty((T, B) + tuple(inner), dtype=torch.float64, device=dev).transpose(0, 1)
    return torch.empty((B, T) + tuple(inner), dtype=torch.float64, device=dev)


def _strides(B, T, time_major):
    return (1, B) if time_major else (T, 1)


def _stream_ptr(stream):
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def _disc_args(disc, B, T, d):
    keep = []
    null = (None, 0)
    if disc.mode == _lib.DISC_GIVEN:
        A, sA = _bview(disc.A, "A", (B, T, d, d), 3)
        Q, sQ = _bview(disc.Q, "Q", (B, T, d, d), 3)
        keep += [A, Q]
        return keep, (A.data_ptr(), sA[0]), (Q.data_ptr(), sQ[0]), null, null
    lam, sl = _bview(disc.lam, "lam", (B, disc.nblk), 1)
    if disc.mode == _lib.DISC_IWP:
        keep += [lam]
        return keep, null, null, (lam.data_ptr(), sl[0]), null
    Pinf, sP = _bview(disc.Pinf, "Pinf", (B, d, d), 2)
    keep += [lam, Pinf]
    return keep, null, null, (lam.data_ptr(), sl[0]), (Pinf.data_ptr(), sP[0])


def kf_wave_series(d, m=1, nblk=0):
    """Series one full wave of the smoother kernel keeps resident on the current GPU for state dim d
    (`physs_kf_wave_series`; nblk > 0: DISC_MATERN with nblk blocks, else DISC_GIVEN).  Batches that are whole
    multiples of it never end on a half-empty GPU."""
    with torch.cuda.device(torch.cuda.current_device()):
        w = int(_lib.load().physs_kf_wave_series(int(d), _lib.DISC_MATERN if nblk else _lib.DISC_GIVEN, int(nblk)))
    if w <= 0:
        raise _lib.PhyssError("physs_kf_wave_series: no fixed-wave kernel for d = %d" % d)
    return w


def spd_inverse(A, jitter=0.0, stream=None):
    """(A + jitter I)^-1 for a batch of small SPD matrices [..., D, D] (D <= 8) -- precision sites R_inv ->
    covariance sites R (`physs_spd_inverse_f64`)."""
    A = _dev(A, "A").contiguous()
    D = A.shape[-1]
    out = torch.empty_like(A)
    N = A.numel() // (D * D)
    with torch.cuda.device(A.device):
        st = _lib.load().physs_spd_inverse_f64(_stream_ptr(stream), N, D, A.data_ptr(), float(jitter), out.data_ptr())
    _lib.check(st, "physs_spd_inverse_f64")
    return out


def kf_supported(d, m, disc):
    return bool(_lib.load().physs_kf_supported(d, m, disc.mode, disc.nblk))


class _Packed:
    """Arguments of the filter / smoother families marshalled for the C ABI (+ tensors kept alive)."""
    pass


def _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream):
    if Y.dim() != 3:
        raise ValueError("Y must be [B, T, m]")
    Y, tmaj = step_layout(Y, "Y")
    B, T, m = Y.shape
    sbs, sts = _strides(B, T, tmaj)
    P0v, sP0 = _bview(P0, "P0", (B, P0.shape[-1], P0.shape[-1]), 2)
    d = P0v.shape[-1]
    m0v, sm0 = _bview(m0, "m0", (B, d), 1)
    dtv, sdt = _bview(dt, "dt", (B, T), 1)
    Rv, sR = _bview(R, "R", (B, T, m, m), 2)   # any batch / time strides (0 = broadcast)
    if H is None:
        Hv, Hptr, sH = None, None, (0,)
    else:
        Hv, sH = _bview(H, "H", (B, m, d), 2)
        Hptr = Hv.data_ptr()
    keep, (pA, bA), (pQ, bQ), (pl, bl), (pPi, bPi) = _disc_args(disc, B, T, d)
    jit = settings.jitter if jitter is None else jitter
    p = _Packed()
    p.B, p.T, p.d, p.m, p.tmaj, p.dev = B, T, d, m, tmaj, Y.device
    p.keep = keep + [Y, P0v, m0v, dtv, Rv, Hv]
    p.head = [_stream_ptr(stream), B, T, sbs, sts, d, m, disc.mode, disc.nblk,
              pA, bA, pQ, bQ, pl, bl, dtv.data_ptr(), sdt[0], pPi, bPi,
              m0v.data_ptr(), sm0[0], P0v.data_ptr(), sP0[0], Hptr, sH[0],
              Y.data_ptr(), Rv.data_ptr(), sR[0], sR[1], float(jit)]
    return p


def _filter_outputs(p, out, want_lml_k):
    B, T, d = p.B, p.T, p.d
    if out is None:
        mf = empty_steps(B, T, (d,), p.dev, p.tmaj)
        Pf = empty_steps(B, T, (d, d), p.dev, p.tmaj)
    else:
        mf, Pf = out
        ok = tuple(mf.shape) == (B, T, d) and tuple(Pf.shape) == (B, T, d, d)
        for o in (mf, Pf):
            ok = ok and (o.transpose(0, 1).is_contiguous() if p.tmaj else o.is_contiguous())
        if not ok:
            raise ValueError("out buffers must be [B,T,d] and [B,T,d,d] in the memory order of Y")
    lml = torch.empty((B,), dtype=torch.float64, device=p.dev)
    lml_k = empty_steps(B, T, (), p.dev, p.tmaj) if want_lml_k else None
    return mf, Pf, lml, lml_k


def kf_filter(dt, Y, R, H, m0, P0, disc, jitter=None, want_lml_k=False, out=None, stream=None):
    """Batched sequential Kalman filter (kalman_filter.py:439-485 semantics per series).

    dt -> [B, T]   (dt[0] = 0, dt[k] = t_k - t_{k-1});  Y [B, T, m] (NaN = missing);
    R -> [B, T, m, m];  H -> [B, m, d] or None (identity, m == d);  m0 -> [B, d];  P0 -> [B, d, d].
    Returns (lml [B], mf [B, T, d], Pf [B, T, d, d][, lml_k [B, T]]).
    `out=(mf, Pf)` reuses caller-provided output buffers.

    Memory order: if Y is a transposed view of a contiguous [T, B, m] tensor (time-major), the outputs
    are produced time-major too (logical shape still [B, T, ...]) -- the fast, coalesced layout.
    """
    lib = _lib.load()
    p = _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream)
    mf, Pf, lml, lml_k = _filter_outputs(p, out, want_lml_k)
    with torch.cuda.device(p.dev):
        st = lib.physs_kf_filter_f64(*p.head, mf.data_ptr(), Pf.data_ptr(), lml.data_ptr(),
                                     lml_k.data_ptr() if lml_k is not None else None)
    _lib.check(st, "physs_kf_filter_f64")
    if want_lml_k:
        return lml, mf, Pf, lml_k
    return lml, mf, Pf


RES_KINDS = {"sin": 0, "cos": 1, "square": 2, "cube": 3, "prod": 4}


def kf_filter_colloc(dt, Y, R, H, m0, P0, disc, res_w, terms=(), forcing=None, y_pseudo=None, boundary=None,
                     observe_data=True, jitter=None, want_lml_k=False, out=None, stream=None):
    """Batched collocation (EKF) filter: kf_predict_step(PDE, 'sequential'), kalman_filter.py:340-427, per series
    (`physs_kf_filter_colloc_f64`).  Arguments as `kf_filter` plus the residual table
        g_p(x, k) = res_w[p] . x + sum_q coef_q phi_q(x[idx_q]) + forcing[p, k]
    res_w [pc, d] (array-like, host); terms: iterable of (output p, kind in RES_KINDS, state index, coefficient) --
    kind "prod" is the bilinear coef * x[i] * x[j] with the index i | (j << 8), or the pair (i, j);
    forcing [pc, T] device tensor or None; y_pseudo [pc] (0 / NaN, default zeros); boundary [B, T, m] device tensor
    (NaN = none) or None.  Returns (lml, mf, Pf[, lml_k])."""
    import ctypes
    lib = _lib.load()
    p = _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream)
    mf, Pf, lml, lml_k = _filter_outputs(p, out, want_lml_k)
    w = np.ascontiguousarray(np.atleast_2d(np.asarray(res_w, np.float64)))
    pc = w.shape[0]
    if w.shape[1] != p.d:
        raise ValueError("res_w must be [pc, d]")
    terms = list(terms)
    t_out = np.array([t[0] for t in terms], np.int32)
    t_kind = np.array([RES_KINDS[t[1]] if isinstance(t[1], str) else int(t[1]) for t in terms], np.int32)
    t_idx = np.array([t[2][0] | (t[2][1] << 8) if isinstance(t[2], (tuple, list)) else t[2] for t in terms], np.int32)
    t_coef = np.array([t[3] for t in terms], np.float64)
    yps = np.zeros(pc) if y_pseudo is None else np.ascontiguousarray(np.asarray(y_pseudo, np.float64).reshape(pc))
    keep = []
    fptr = None
    if forcing is not None:
        f = _dev(forcing, "forcing").contiguous()
        if tuple(f.shape) != (pc, p.T):
            raise ValueError("forcing must be [pc, T]")
        keep.append(f)
        fptr = f.data_ptr()
    bptr = None
    if boundary is not None:
        bd, btm = step_layout(boundary, "boundary")
        if tuple(bd.shape) != (p.B, p.T, p.m) or btm != p.tmaj:
            if tuple(bd.shape) != (p.B, p.T, p.m):
                raise ValueError("boundary must be [B, T, m]")
            bd = bd.transpose(0, 1).contiguous().transpose(0, 1) if p.tmaj else bd.contiguous()
        keep.append(bd)
        bptr = bd.data_ptr()

    def hp(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a.size else None
    with torch.cuda.device(p.dev):
        st = lib.physs_kf_filter_colloc_f64(*p.head, pc, hp(w), len(terms), hp(t_out), hp(t_kind), hp(t_idx),
                                            hp(t_coef), fptr, hp(yps), bptr, 1 if observe_data else 0,
                                            mf.data_ptr(), Pf.data_ptr(), lml.data_ptr(),
                                            lml_k.data_ptr() if lml_k is not None else None)
    _lib.check(st, "physs_kf_filter_colloc_f64")
    if want_lml_k:
        return lml, mf, Pf, lml_k
    return lml, mf, Pf


def kf_filter_smooth(dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s=None, Hout=None, jitter=None, stream=None):
    """Filter + smoother in ONE C-ABI call (physs_kf_filter_smooth_f64): same results as `kf_filter` followed by
    `rts_smooth`.  dt_f / dt_s are the two dt conventions; disc_s (DISC_GIVEN only) carries the smoother's A_k, Q_k.
    Returns (lml, mf, Pf, ms, Ps)."""
    lib = _lib.load()
    p = _pack_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter, stream)
    mf, Pf, lml, _ = _filter_outputs(p, None, False)
    B, T, d = p.B, p.T, p.d
    dtv, sdt = _bview(dt_s, "dt_s", (B, T), 1)
    pA = pQ = None
    keep = []
    if disc_f.mode == _lib.DISC_GIVEN:
        if disc_s is None:
            raise ValueError("DISC_GIVEN needs disc_s (the smoother's transitions)")
        keep, (pA, bA), (pQ, bQ), _, _ = _disc_args(disc_s, B, T, d)
        if (bA, bQ) != (p.head[10], p.head[12]):
            raise ValueError("disc_f and disc_s must share their batch layout")
    if Hout is None:
        mo, Hptr, mp = 0, None, d
    else:
        Hout = _dev(Hout, "Hout").contiguous()
        mo, Hptr, mp = Hout.shape[0], Hout.data_ptr(), Hout.shape[0]
    ms = empty_steps(B, T, (mp,), p.dev, p.tmaj)
    Ps = empty_steps(B, T, (mp, mp), p.dev, p.tmaj)
    with torch.cuda.device(p.dev):
        st = lib.physs_kf_filter_smooth_f64(*p.head, pA, pQ, dtv.data_ptr(), sdt[0], Hptr, mo, mf.data_ptr(),
                                            Pf.data_ptr(), lml.data_ptr(), None, ms.data_ptr(), Ps.data_ptr())
    _lib.check(st, "physs_kf_filter_smooth_f64")
    del keep
    return lml, mf, Pf, ms, Ps


def kf_filter_smooth_packed_supported(Y, d, m, disc):
    """True when `kf_filter_smooth_packed` covers this problem: register kernels (d <= 4) and Y [B, T, m]
    time-major (a transposed view of a contiguous [T, B, m] tensor)."""
    lib = _lib.load()
    if Y.dim() != 3 or not lib.physs_kf_filter_smooth_packed_supported(d, m, disc.mode, disc.nblk):
        return False
    return bool(step_layout(Y, "Y")[1])


def kf_filter_smooth_packed(dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s=None, Hout=None, jitter=None, ws=None,
                            stream=None):
    """Filter + smoother in ONE C-ABI call that never writes the filtered moments as an output
    (physs_kf_filter_smooth_packed_f64): the hand-over between the two passes is a workspace of packed rows
    [m | triu(P)].  Same arguments as `kf_filter_smooth`; returns (lml, ms, Ps): lml and full-state (ms, Ps) bitwise
    what `kf_filter_smooth` returns, projected outputs to 1e-14.  `ws`: optional caller-kept uint8 workspace (reused across calls when large enough)."""
    lib = _lib.load()
    p = _pack_filter(dt_f, Y, R, H, m0, P0, disc_f, jitter, stream)
    B, T, d = p.B, p.T, p.d
    if not p.tmaj:
        raise NotImplementedError("kf_filter_smooth_packed: time-major steps only (see kf_filter_smooth_packed_supported)")
    dtv, sdt = _bview(dt_s, "dt_s", (B, T), 1)
    pA = pQ = None
    keep = []
    if disc_f.mode == _lib.DISC_GIVEN:
        if disc_s is None:
            raise ValueError("DISC_GIVEN needs disc_s (the smoother's transitions)")
        keep, (pA, bA), (pQ, bQ), _, _ = _disc_args(disc_s, B, T, d)
        if (bA, bQ) != (p.head[10], p.head[12]):
            raise ValueError("disc_f and disc_s must share their batch layout")
    if Hout is None:
        mo, Hptr, mp = 0, None, d
    else:
        Hout = _dev(Hout, "Hout").contiguous()
        mo, Hptr, mp = Hout.shape[0], Hout.data_ptr(), Hout.shape[0]
    need = int(lib.physs_kf_filter_smooth_packed_ws_bytes(B, T, p.head[4], d))
    if ws is None or ws.numel() < need or ws.device != p.dev:
        ws = torch.empty((need,), dtype=torch.uint8, device=p.dev)
    lml = torch.empty((B,), dtype=torch.float64, device=p.dev)
    ms = empty_steps(B, T, (mp,), p.dev, True)
    Ps = empty_steps(B, T, (mp, mp), p.dev, True)
    with torch.cuda.device(p.dev):
        st = lib.physs_kf_filter_smooth_packed_f64(*p.head, pA, pQ, dtv.data_ptr(), sdt[0], Hptr, mo, ws.data_ptr(),
                                                   ws.numel(), lml.data_ptr(), None, ms.data_ptr(), Ps.data_ptr())
    _lib.check(st, "physs_kf_filter_smooth_packed_f64")
    del keep
    return lml, ms, Ps


def kf_filter_vjp(dt, Y, R, H, m0, P0, disc, mf, Pf, g_lml=None, jitter=None, want_R_step=False, stream=None):
    """Reverse pass of `kf_filter`'s lml (include/physs_b200.h: physs_kf_filter_vjp_f64): gradients of
    sum_b g_lml[b] * lml[b] with respect to the filter's inputs, given its outputs (mf, Pf).

    Returns a dict: DISC_GIVEN -> 'gA', 'gQ' [B, T, d, d]; DISC_MATERN -> 'glam' [B, nblk], 'gPinf' [B, d, d];
    always 'gH' [B, m, d], 'gR' [B, m, m] (summed over the steps), 'gm0' [B, d], 'gP0' [B, d, d];
    'gR_step' [B, T, m, m] with want_R_step.  d <= 4, m == 1: register kernel, any discretisation; otherwise
    d <= 32, m <= d with Disc.given (lane-group kernel; chain (gA, gQ) to hyper-parameters with torch, see
    models.SDE_GP.log_marginal_likelihood_and_grad)."""
    lib = _lib.load()
    p = _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream)
    B, T, d, m, dev = p.B, p.T, p.d, p.m, p.dev
    if not lib.physs_kf_vjp_supported(d, m, disc.mode, disc.nblk):
        raise NotImplementedError("kf_filter_vjp: d <= 4 with m == 1 (any discretisation) or d <= 32, m <= d with "
                                  "Disc.given (d=%d, m=%d)" % (d, m))
    mfv, tm1 = step_layout(_dev(mf, "mf"), "mf")
    Pfv, tm2 = step_layout(_dev(Pf, "Pf"), "Pf")
    if tm1 != p.tmaj or tm2 != p.tmaj:
        raise ValueError("kf_filter_vjp: mf / Pf must be in the memory order of Y (the filter's own outputs are)")
    z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)      # noqa: E731
    out = {"gH": z(B, m, d), "gR": z(B, m, m), "gm0": z(B, d), "gP0": z(B, d, d)}
    gA = gQ = glam = gPinf = gRs = None
    if disc.mode == _lib.DISC_GIVEN:
        gA = empty_steps(B, T, (d, d), dev, p.tmaj)
        gQ = empty_steps(B, T, (d, d), dev, p.tmaj)
        out["gA"], out["gQ"] = gA, gQ
    else:
        glam, gPinf = z(B, disc.nblk), z(B, d, d)
        out["glam"], out["gPinf"] = glam, gPinf
    if want_R_step:
        gRs = empty_steps(B, T, (m, m), dev, p.tmaj)
        out["gR_step"] = gRs
    g = None
    if g_lml is not None:
        g = _dev(g_lml, "g_lml").reshape(-1).contiguous()
        if g.numel() != B:
            raise ValueError("g_lml must have one entry per series")
    ptr = lambda t: t.data_ptr() if t is not None else None                      # noqa: E731
    with torch.cuda.device(dev):
        st = lib.physs_kf_filter_vjp_f64(*p.head, mfv.data_ptr(), Pfv.data_ptr(), ptr(g), ptr(gA), ptr(gQ),
                                         ptr(glam), ptr(gPinf), out["gH"].data_ptr(), ptr(gRs),
                                         out["gR"].data_ptr(), out["gm0"].data_ptr(), out["gP0"].data_ptr())
    _lib.check(st, "physs_kf_filter_vjp_f64")
    return out


def _pack_smooth(dt, mf, Pf, disc, Hout, jitter, stream):
    mf, tmaj = step_layout(mf, "mf")
    Pf, tmaj_P = step_layout(Pf, "Pf")
    if tmaj != tmaj_P:
        mf, Pf, tmaj = mf.contiguous(), Pf.contiguous(), False
    B, T, d = mf.shape
    sbs, sts = _strides(B, T, tmaj)
    dtv, sdt = _bview(dt, "dt", (B, T), 1)
    keep, (pA, bA), (pQ, bQ), (pl, bl), (pPi, bPi) = _disc_args(disc, B, T, d)
    if Hout is None:
        mo, Hptr, mp = 0, None, d
    else:
        Hout = _dev(Hout, "Hout").contiguous()
        mo = Hout.shape[0]
        Hptr, mp = Hout.data_ptr(), mo
    jit = settings.jitter if jitter is None else jitter
    p = _Packed()
    p.B, p.T, p.d, p.mp, p.tmaj, p.dev = B, T, d, mp, tmaj, mf.device
    p.keep = keep + [mf, Pf, dtv, Hout]
    p.head = [_stream_ptr(stream), B, T, sbs, sts, d, disc.mode, disc.nblk,
              pA, bA, pQ, bQ, pl, bl, dtv.data_ptr(), sdt[0], pPi, bPi,
              mf.data_ptr(), Pf.data_ptr(), Hptr, mo, float(jit)]
    return p


def _smooth_outputs(p, out):
    if out is None:
        ms = empty_steps(p.B, p.T, (p.mp,), p.dev, p.tmaj)
        Ps = empty_steps(p.B, p.T, (p.mp, p.mp), p.dev, p.tmaj)
    else:
        ms, Ps = out
        for o in (ms, Ps):
            if not (o.transpose(0, 1).is_contiguous() if p.tmaj else o.is_contiguous()):
                raise ValueError("out buffers must be in the memory order of mf / Pf")
    return ms, Ps


def rts_smooth(dt, mf, Pf, disc, Hout=None, jitter=None, out=None, stream=None):
    """Batched sequential RTS smoother (rts_smoother.py:162-192 semantics per series).

    dt -> [B, T] (dt[k] = t_{k+1} - t_k, dt[T-1] = 0);  mf [B, T, d], Pf [B, T, d, d];
    Hout [mo, d] projects the output (None = full_state=True).
    Returns (ms [B, T, mo'], Ps [B, T, mo', mo']).
    """
    lib = _lib.load()
    p = _pack_smooth(dt, mf, Pf, disc, Hout, jitter, stream)
    ms, Ps = _smooth_outputs(p, out)
    with torch.cuda.device(p.dev):
        st = lib.physs_rts_smooth_f64(*p.head, ms.data_ptr(), Ps.data_ptr())
    _lib.check(st, "physs_rts_smooth_f64")
    return ms, Ps


# ------------------------------------------------------------------------------- parallel-in-time
def default_chunk_len(B, T, d=8):
    """Chunk length that gives the GPU enough independent (series, chunk) pairs, clamped to [128, T]:
    ~64k for the register kernels (d <= 4: one THREAD per pair), ~148 SMs x 64 lane groups above.
    Long single series get many short chunks, large batches few long ones."""
    target = 65536 if d <= 4 else 148 * 64
    want_chunks = max(1, (target + B - 1) // B)
    # not below 128 steps: the fix-up passes contract the O(jitter) boundary error by the filter's forgetting
    # over ONE chunk, so chunks shorter than the mixing time cost more passes than they gain in parallelism
    return even_chunk_len(T, int(min(T, max(128, -(-T // want_chunks)))))


def even_chunk_len(T, L, slack=0.25):
    """The chunk length closest to L (within +-slack) that divides T, else L.  Without a ragged last chunk
    every pass of the scan is ONE launch per kernel: the ragged chunk runs as a launch of its own (its
    lane groups take fewer steps than the others of a warp would), serialised behind the main one."""
    T, L = int(T), int(L)
    if L <= 0 or T <= L or T % L == 0:
        return L
    best = None
    for c in range(max(1, int(L * (1 - slack))), int(L * (1 + slack)) + 1):
        if T % c == 0 and (best is None or abs(c - L) < abs(best - L)):
            best = c
    return best if best is not None else L


def pscan_workspace(B, T, d, chunk_len, dev):
    n = _lib.load().physs_pscan_workspace_bytes(B, T, d, chunk_len)
    return torch.empty((n // 8,), dtype=torch.float64, device=dev)


def _polish_default(jit, polish):
    """Fix-up passes after the scan.  Each pass contracts the O(jitter) boundary error by the filter's
    forgetting over one chunk; a pass whose chunks already agree exits after `patience` steps, so spare
    passes cost a few steps per chunk.  None: 4 passes when jitter != 0, none when the scan is exact."""
    if polish is None:
        return 4 if jit != 0.0 else 0
    return int(polish)


def pscan_filter(dt, Y, R, H, m0, P0, disc, chunk_len=None, jitter=None, polish=None, delta=1e-10, patience=4,
                 want_lml_k=False, out=None, ws=None, stream=None, return_status=False):
    """Parallel-in-time Kalman filter: same arguments / results as kf_filter (see include/physs_b200.h,
    'Parallel-in-time forms').  `polish=None` picks 4 fix-up passes when jitter != 0 and none otherwise.
    With return_status the device flag (1 = a chunk did not converge during polishing) is appended."""
    lib = _lib.load()
    p = _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream)
    L = default_chunk_len(p.B, p.T, p.d) if chunk_len is None else int(chunk_len)
    mf, Pf, lml, lml_k = _filter_outputs(p, out, want_lml_k)
    ws = pscan_workspace(p.B, p.T, p.d, L, p.dev) if ws is None else ws
    status = torch.zeros((1,), dtype=torch.int32, device=p.dev)
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_filter_f64(*p.head, L, _polish_default(p.head[-1], polish), float(delta), int(patience),
                                        ws.data_ptr(), mf.data_ptr(), Pf.data_ptr(), lml.data_ptr(),
                                        lml_k.data_ptr() if lml_k is not None else None, status.data_ptr())
    _lib.check(st, "physs_pscan_filter_f64")
    res = (lml, mf, Pf) + ((lml_k,) if want_lml_k else ())
    return res + ((status,) if return_status else ())


def pscan_filter_spec(dt, Y, R, H, m0, P0, disc, chunk_len=None, warm=None, jitter=None, polish=2, delta=1e-11,
                      patience=4, want_lml_k=False, out=None, ws=None, stream=None):
    """Speculative parallel-in-time filter (include/physs_b200.h): warm-up instead of summaries + scan, verified
    by the fix-up passes.  Returns (lml, mf, Pf[, lml_k], status); status = 1 -> use pscan_filter."""
    lib = _lib.load()
    p = _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream)
    L = default_chunk_len(p.B, p.T, p.d) if chunk_len is None else int(chunk_len)
    W = min(L, 128) if warm is None else int(warm)
    mf, Pf, lml, lml_k = _filter_outputs(p, out, want_lml_k)
    ws = pscan_workspace(p.B, p.T, p.d, L, p.dev) if ws is None else ws
    status = torch.zeros((1,), dtype=torch.int32, device=p.dev)
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_filter_spec_f64(*p.head, L, W, int(polish), float(delta), int(patience), ws.data_ptr(),
                                             mf.data_ptr(), Pf.data_ptr(), lml.data_ptr(),
                                             lml_k.data_ptr() if lml_k is not None else None, status.data_ptr())
    _lib.check(st, "physs_pscan_filter_spec_f64")
    return (lml, mf, Pf) + ((lml_k,) if want_lml_k else ()) + (status,)


def pscan_smooth_spec(dt, mf, Pf, disc, chunk_len=None, warm=None, jitter=None, polish=2, delta=1e-11, patience=4,
                      out=None, ws=None, stream=None):
    """Speculative parallel-in-time smoother (full-state output).  Returns (ms, Ps, status)."""
    lib = _lib.load()
    p = _pack_smooth(dt, mf, Pf, disc, None, jitter, stream)
    L = default_chunk_len(p.B, p.T, p.d) if chunk_len is None else int(chunk_len)
    W = min(L, 128) if warm is None else int(warm)
    ms, Ps = _smooth_outputs(p, out)
    ws = pscan_workspace(p.B, p.T, p.d, L, p.dev) if ws is None else ws
    status = torch.zeros((1,), dtype=torch.int32, device=p.dev)
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_smooth_spec_f64(*p.head, L, W, int(polish), float(delta), int(patience), ws.data_ptr(),
                                             ms.data_ptr(), Ps.data_ptr(), status.data_ptr())
    _lib.check(st, "physs_pscan_smooth_spec_f64")
    return ms, Ps, status


def pscan_smooth(dt, mf, Pf, disc, Hout=None, chunk_len=None, jitter=None, out=None, ws=None, stream=None):
    """Parallel-in-time RTS smoother: same arguments / results as rts_smooth."""
    lib = _lib.load()
    p = _pack_smooth(dt, mf, Pf, disc, Hout, jitter, stream)
    L = default_chunk_len(p.B, p.T, p.d) if chunk_len is None else int(chunk_len)
    ms, Ps = _smooth_outputs(p, out)
    ws = pscan_workspace(p.B, p.T, p.d, L, p.dev) if ws is None else ws
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_smooth_f64(*p.head, L, ws.data_ptr(), ms.data_ptr(), Ps.data_ptr())
    _lib.check(st, "physs_pscan_smooth_f64")
    return ms, Ps


# ---- time-sharded building blocks (used by physs_gp_b200/timeshard.py)
def pscan_filter_local(dt, Y, R, H, m0, P0, disc, chunk_len, ws, jitter=None, stream=None, out=None):
    """Scan element of this whole time range, [B, 3 d^2 + 2 d]; prefixes stay in `ws` for *_finish.
    `out`: a contiguous [B, 3 d^2 + 2 d] view to write it into -- the time-sharded driver passes this rank's slot
    of the all-gather buffer, so the summary kernel's store IS the collective's send buffer."""
    lib = _lib.load()
    p = _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream)
    ne = 3 * p.d * p.d + 2 * p.d
    if out is not None and (tuple(out.shape) != (p.B, ne) or not out.is_contiguous() or out.dtype != torch.float64):
        raise ValueError("pscan_filter_local: out must be a contiguous float64 [B, 3 d^2 + 2 d] tensor")
    total = out if out is not None else torch.empty((p.B, ne), dtype=torch.float64, device=p.dev)
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_filter_local_f64(*p.head, int(chunk_len), ws.data_ptr(), total.data_ptr())
    _lib.check(st, "physs_pscan_filter_local_f64")
    return total


def pscan_filter_fold(totals, m0, P0, stream=None):
    """(m0, P0) [B, d], [B, d, d] pushed through totals [K, B, ne] in time order."""
    lib = _lib.load()
    K, B = totals.shape[0], totals.shape[1]
    d = P0.shape[-1]
    m0v, sm0 = _bview(m0, "m0", (B, d), 1)
    P0v, sP0 = _bview(P0, "P0", (B, d, d), 2)
    totals = _dev(totals, "totals").contiguous()
    mo = torch.empty((B, d), dtype=torch.float64, device=totals.device)
    Po = torch.empty((B, d, d), dtype=torch.float64, device=totals.device)
    with torch.cuda.device(totals.device):
        st = lib.physs_pscan_filter_fold_f64(_stream_ptr(stream), B, d, K, totals.data_ptr(), m0v.data_ptr(), sm0[0],
                                             P0v.data_ptr(), sP0[0], mo.data_ptr(), Po.data_ptr())
    _lib.check(st, "physs_pscan_filter_fold_f64")
    return mo, Po


def pscan_filter_finish(dt, Y, R, H, m0, P0, disc, chunk_len, ws, start=None, jitter=None, polish=None, delta=1e-10,
                        patience=4, want_lml_k=False, out=None, stream=None):
    lib = _lib.load()
    p = _pack_filter(dt, Y, R, H, m0, P0, disc, jitter, stream)
    mf, Pf, lml, lml_k = _filter_outputs(p, out, want_lml_k)
    status = torch.zeros((1,), dtype=torch.int32, device=p.dev)
    sm = sP = None
    if start is not None:
        sm, sP = _dev(start[0], "start_m").contiguous(), _dev(start[1], "start_P").contiguous()
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_filter_finish_f64(*p.head, int(chunk_len), _polish_default(p.head[-1], polish),
                                               float(delta), int(patience), ws.data_ptr(),
                                               None if sm is None else sm.data_ptr(),
                                               None if sP is None else sP.data_ptr(),
                                               mf.data_ptr(), Pf.data_ptr(), lml.data_ptr(),
                                               lml_k.data_ptr() if lml_k is not None else None, status.data_ptr())
    _lib.check(st, "physs_pscan_filter_finish_f64")
    return (lml, mf, Pf) + ((lml_k,) if want_lml_k else ()) + (status,)


def pscan_smooth_local(dt, mf, Pf, disc, chunk_len, ws, jitter=None, stream=None):
    lib = _lib.load()
    p = _pack_smooth(dt, mf, Pf, disc, None, jitter, stream)
    total = torch.empty((p.B, 2 * p.d * p.d + p.d), dtype=torch.float64, device=p.dev)
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_smooth_local_f64(*p.head, int(chunk_len), ws.data_ptr(), total.data_ptr())
    _lib.check(st, "physs_pscan_smooth_local_f64")
    return total


def pscan_smooth_fold(totals, m_end, P_end, stream=None):
    """(m_end, P_end) [B, d], [B, d, d] pulled back through totals [K, B, ns] (time order, last first)."""
    lib = _lib.load()
    K, B = totals.shape[0], totals.shape[1]
    d = P_end.shape[-1]
    totals = _dev(totals, "totals").contiguous()
    m_end, P_end = _dev(m_end, "m_end").contiguous(), _dev(P_end, "P_end").contiguous()
    mo = torch.empty((B, d), dtype=torch.float64, device=totals.device)
    Po = torch.empty((B, d, d), dtype=torch.float64, device=totals.device)
    with torch.cuda.device(totals.device):
        st = lib.physs_pscan_smooth_fold_f64(_stream_ptr(stream), B, d, K, totals.data_ptr(), m_end.data_ptr(),
                                             P_end.data_ptr(), mo.data_ptr(), Po.data_ptr())
    _lib.check(st, "physs_pscan_smooth_fold_f64")
    return mo, Po


def pscan_smooth_finish(dt, mf, Pf, disc, chunk_len, ws, start=None, Hout=None, jitter=None, out=None, stream=None):
    lib = _lib.load()
    p = _pack_smooth(dt, mf, Pf, disc, Hout, jitter, stream)
    ms, Ps = _smooth_outputs(p, out)
    sm = sP = None
    if start is not None:
        sm, sP = _dev(start[0], "start_m").contiguous(), _dev(start[1], "start_P").contiguous()
    with torch.cuda.device(p.dev):
        st = lib.physs_pscan_smooth_finish_f64(*p.head, int(chunk_len), ws.data_ptr(),
                                               None if sm is None else sm.data_ptr(),
                                               None if sP is None else sP.data_ptr(), ms.data_ptr(), Ps.data_ptr())
    _lib.check(st, "physs_pscan_smooth_finish_f64")
    return ms, Ps


def fp64_peak_tflops(dev=None, iters=20000):
    """Measured DFMA throughput of this GPU in TFLOP/s (CUDA events around physs_fp64_probe)."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else dev
    out = torch.zeros((1,), dtype=torch.float64, device=dev)
    blocks = 148 * 8
    best = 0.0
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream()
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.physs_fp64_probe(st.cuda_stream, blocks, iters, out.data_ptr()), "physs_fp64_probe")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, blocks * 256 * iters * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


# ------------------------------------------------------------------------------------- large blocks
class BigDisc:
    """Discretisation of the large-block path: A, Q [nA, d, d] device tensors, one pair per DISTINCT step
    size, and a host int32 index [T] selecting the pair of every step."""

    def __init__(self, A, Q, index):
        import numpy as np
        self.A, self.Q = _dev(A, "A").contiguous(), _dev(Q, "Q").contiguous()
        self.index = np.ascontiguousarray(index, dtype=np.int32)


def _big_ws(d, m, dev):
    n = _lib.load_big().physs_big_workspace_bytes(d, m)
    if n <= 0:
        raise _lib.PhyssError("physs_big_workspace_bytes failed (cuSOLVER handle?)")
    return torch.empty((n // 8 + 2,), dtype=torch.float64, device=dev)


def kf_filter_big(Y, R, H, m0, P0, disc, jitter=None, stream=None):
    """One series, large state: Y [T, m], R [T|1, m, m], H [m, d], m0 [d], P0 [d, d] -> (lml [], mf, Pf)."""
    import ctypes
    lib = _lib.load_big()
    Y, H, m0, P0 = (_dev(x, n).contiguous() for x, n in ((Y, "Y"), (H, "H"), (m0, "m0"), (P0, "P0")))
    R = _dev(R, "R").contiguous()
    T, m = Y.shape
    d = P0.shape[-1]
    R_ts = m * m if (R.dim() == 3 and R.shape[0] > 1) else 0
    dev = Y.device
    mf = torch.empty((T, d), dtype=torch.float64, device=dev)
    Pf = torch.empty((T, d, d), dtype=torch.float64, device=dev)
    lml = torch.empty((1,), dtype=torch.float64, device=dev)
    ws = _big_ws(d, m, dev)
    jit = settings.jitter if jitter is None else jitter
    idx = disc.index.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    with torch.cuda.device(dev):
        st = lib.physs_kf_filter_big_f64(_stream_ptr(stream), T, d, m, disc.A.data_ptr(), disc.Q.data_ptr(), idx,
                                         m0.data_ptr(), P0.data_ptr(), H.data_ptr(), Y.data_ptr(), R.data_ptr(),
                                         R_ts, float(jit), ws.data_ptr(), ws.numel() * 8, mf.data_ptr(),
                                         Pf.data_ptr(), lml.data_ptr())
    _lib.check_big(st, "physs_kf_filter_big_f64")
    return lml[0], mf, Pf


def rts_smooth_big(mf, Pf, disc, Hout=None, jitter=None, stream=None):
    import ctypes
    lib = _lib.load_big()
    mf, Pf = _dev(mf, "mf").contiguous(), _dev(Pf, "Pf").contiguous()
    T, d = mf.shape
    dev = mf.device
    if Hout is None:
        mo, Hp, mp = 0, None, d
    else:
        Hout = _dev(Hout, "Hout").contiguous()
        mo, Hp, mp = Hout.shape[0], Hout.data_ptr(), Hout.shape[0]
    ms = torch.empty((T, mp), dtype=torch.float64, device=dev)
    Ps = torch.empty((T, mp, mp), dtype=torch.float64, device=dev)
    ws = _big_ws(d, d, dev)
    jit = settings.jitter if jitter is None else jitter
    idx = disc.index.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    with torch.cuda.device(dev):
        st = lib.physs_rts_smooth_big_f64(_stream_ptr(stream), T, d, disc.A.data_ptr(), disc.Q.data_ptr(), idx,
                                          mf.data_ptr(), Pf.data_ptr(), Hp, mo, float(jit), ws.data_ptr(),
                                          ws.numel() * 8, ms.data_ptr(), Ps.data_ptr())
    _lib.check_big(st, "physs_rts_smooth_big_f64")
    return ms, Ps


# ------------------------------------------------------- separable spatio-temporal prior, hand-written kernels
kron_prof = {}       # PHYSS_KRON_PROF=1: device views of the phase timers of the last call (measurement aid)


class KronDisc:
    """Discretisation of the separable spatio-temporal route (include/physs_b200.h, physs_kf_filter_kron_f64):
    temporal At, Qt [nA, ds, ds] per DISTINCT step size, a DEVICE int32 index [T] selecting the pair of every
    step, and the spatial Gram matrix Ks [Ns, Ns]."""

    def __init__(self, At, Qt, index, Ks):
        self.At, self.Qt = _dev(At, "At").contiguous(), _dev(Qt, "Qt").contiguous()
        self.Ks = _dev(Ks, "Ks").contiguous()
        self.index = index.to(device=self.At.device, dtype=torch.int32).contiguous()
        self.ds = int(self.At.shape[-1])
        self.Ns = int(self.Ks.shape[-1])


def _kron_ws(T, Ns, ds, smoother, dev):
    with torch.cuda.device(dev):
        n = _lib.load().physs_kron_workspace_bytes(T, Ns, ds, 1 if smoother else 0)
    if n <= 0:
        raise _lib.PhyssError("physs_kron_workspace_bytes failed: " + (_lib.load().physs_last_error() or b"").decode())
    return torch.empty((n // 8 + 2,), dtype=torch.float64, device=dev)


def kf_filter_kron(Y, R, m0, P0, disc, jitter=None, stream=None):
    """One series, separable prior: Y [T, Ns], R [T|1, Ns, Ns], m0 [d], P0 [d, d] -> (lml [], mf [T, d], Pf [T, d, d])."""
    lib = _lib.load()
    Y, m0, P0 = (_dev(x, n).contiguous() for x, n in ((Y, "Y"), (m0, "m0"), (P0, "P0")))
    R = _dev(R, "R").contiguous()
    T, Ns = Y.shape
    ds, d = disc.ds, disc.ds * disc.Ns
    if Ns != disc.Ns or P0.shape[-1] != d or R.shape[-1] != Ns:
        raise ValueError("kf_filter_kron: inconsistent shapes")
    R_ts = Ns * Ns if (R.dim() == 3 and R.shape[0] > 1) else 0
    dev = Y.device
    mf = torch.empty((T, d), dtype=torch.float64, device=dev)
    Pf = torch.empty((T, d, d), dtype=torch.float64, device=dev)
    lml = torch.empty((1,), dtype=torch.float64, device=dev)
    ws = _kron_ws(T, Ns, ds, False, dev)
    jit = settings.jitter if jitter is None else jitter
    with torch.cuda.device(dev):
        st = lib.physs_kf_filter_kron_f64(_stream_ptr(stream), T, Ns, ds, disc.At.data_ptr(), disc.Qt.data_ptr(),
                                          disc.index.data_ptr(), disc.Ks.data_ptr(), m0.data_ptr(), P0.data_ptr(),
                                          Y.data_ptr(), R.data_ptr(), R_ts, float(jit), ws.data_ptr(),
                                          ws.numel() * 8, mf.data_ptr(), Pf.data_ptr(), lml.data_ptr())
    _lib.check(st, "physs_kf_filter_kron_f64")
    if os.environ.get("PHYSS_KRON_PROF"):
        o = lib.physs_kron_prof_offset(T, Ns, ds, 0)
        kron_prof["filter"] = ws[o:o + 32]
    return lml[0], mf, Pf


def rts_smooth_kron(mf, Pf, disc, project=False, jitter=None, stream=None):
    """project=False -> (ms [T, d], Ps [T, d, d]); True -> (H ms [T, Ns], H Ps H^T [T, Ns, Ns])."""
    lib = _lib.load()
    mf, Pf = _dev(mf, "mf").contiguous(), _dev(Pf, "Pf").contiguous()
    T, d = mf.shape
    ds, Ns = disc.ds, disc.Ns
    if d != ds * Ns:
        raise ValueError("rts_smooth_kron: inconsistent shapes")
    dev = mf.device
    mo = Ns if project else d
    ms = torch.empty((T, mo), dtype=torch.float64, device=dev)
    Ps = torch.empty((T, mo, mo), dtype=torch.float64, device=dev)
    ws = _kron_ws(T, Ns, ds, True, dev)
    jit = settings.jitter if jitter is None else jitter
    with torch.cuda.device(dev):
        st = lib.physs_rts_smooth_kron_f64(_stream_ptr(stream), T, Ns, ds, disc.At.data_ptr(), disc.Qt.data_ptr(),
                                           disc.index.data_ptr(), disc.Ks.data_ptr(), mf.data_ptr(), Pf.data_ptr(),
                                           1 if project else 0, float(jit), ws.data_ptr(), ws.numel() * 8,
                                           ms.data_ptr(), Ps.data_ptr())
    _lib.check(st, "physs_rts_smooth_kron_f64")
    if os.environ.get("PHYSS_KRON_PROF"):
        o = lib.physs_kron_prof_offset(T, Ns, ds, 1)
        kron_prof["smoother"] = ws[o:o + 32 + 2048]
    return ms, Ps
