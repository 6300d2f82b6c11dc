"""Device-level entry points: thin, allocation-explicit wrappers over the C ABI
(include/physs_b200.h) on torch CUDA tensors.  torch is plumbing here (device memory, streams);
every FLOP of the path runs in libphyss_b200.so.  No CPU fallback: non-CUDA tensors raise.

Broadcasting: every argument is viewed with `expand` to its full [B, ...] shape and the resulting
batch (and, for R, time) stride is handed to the kernel; a stride of 0 means "shared".
"""
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
        v = x.contiguous().expand(*shape)
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
    Pinf, sP = _bview(disc.Pinf, "Pinf", (B, d, d), 2)
    keep += [lam, Pinf]
    return keep, null, null, (lam.data_ptr(), sl[0]), (Pinf.data_ptr(), sP[0])


def kf_supported(d, m, disc):
    return bool(_lib.load().physs_kf_supported(d, m, disc.mode, disc.nblk))


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
        Hptr, sH = None, (0,)
    else:
        Hv, sH = _bview(H, "H", (B, m, d), 2)
        Hptr = Hv.data_ptr()
    keep, (pA, bA), (pQ, bQ), (pl, bl), (pPi, bPi) = _disc_args(disc, B, T, d)
    dev = Y.device
    if out is None:
        mf = empty_steps(B, T, (d,), dev, tmaj)
        Pf = empty_steps(B, T, (d, d), dev, tmaj)
    else:
        mf, Pf = out
        ok = tuple(mf.shape) == (B, T, d) and tuple(Pf.shape) == (B, T, d, d)
        for o in (mf, Pf):
            ok = ok and (o.transpose(0, 1).is_contiguous() if tmaj else o.is_contiguous())
        if not ok:
            raise ValueError("out buffers must be [B,T,d] and [B,T,d,d] in the memory order of Y")
    lml = torch.empty((B,), dtype=torch.float64, device=dev)
    lml_k = empty_steps(B, T, (), dev, tmaj) if want_lml_k else None
    jit = settings.jitter if jitter is None else jitter
    with torch.cuda.device(dev):
        st = lib.physs_kf_filter_f64(
            _stream_ptr(stream), B, T, sbs, sts, d, m, disc.mode, disc.nblk,
            pA, bA, pQ, bQ, pl, bl, dtv.data_ptr(), sdt[0], pPi, bPi,
            m0v.data_ptr(), sm0[0], P0v.data_ptr(), sP0[0], Hptr, sH[0],
            Y.data_ptr(), Rv.data_ptr(), sR[0], sR[1], float(jit),
            mf.data_ptr(), Pf.data_ptr(), lml.data_ptr(),
            lml_k.data_ptr() if lml_k is not None else None)
    _lib.check(st, "physs_kf_filter_f64")
    if want_lml_k:
        return lml, mf, Pf, lml_k
    return lml, mf, Pf


def rts_smooth(dt, mf, Pf, disc, Hout=None, jitter=None, out=None, stream=None):
    """Batched sequential RTS smoother (rts_smoother.py:162-192 semantics per series).

    dt -> [B, T] (dt[k] = t_{k+1} - t_k, dt[T-1] = 0);  mf [B, T, d], Pf [B, T, d, d];
    Hout [mo, d] projects the output (None = full_state=True).
    Returns (ms [B, T, mo'], Ps [B, T, mo', mo']).
    """
    lib = _lib.load()
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
    dev = mf.device
    if out is None:
        ms = empty_steps(B, T, (mp,), dev, tmaj)
        Ps = empty_steps(B, T, (mp, mp), dev, tmaj)
    else:
        ms, Ps = out
        for o in (ms, Ps):
            if not (o.transpose(0, 1).is_contiguous() if tmaj else o.is_contiguous()):
                raise ValueError("out buffers must be in the memory order of mf / Pf")
    jit = settings.jitter if jitter is None else jitter
    with torch.cuda.device(dev):
        st = lib.physs_rts_smooth_f64(
            _stream_ptr(stream), B, T, sbs, sts, d, disc.mode, disc.nblk,
            pA, bA, pQ, bQ, pl, bl, dtv.data_ptr(), sdt[0], pPi, bPi,
            mf.data_ptr(), Pf.data_ptr(), Hptr, mo, float(jit), ms.data_ptr(), Ps.data_ptr())
    _lib.check(st, "physs_rts_smooth_f64")
    return ms, Ps
