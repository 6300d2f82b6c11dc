"""The drop-in operators: `filter` / `smoother` registered under filter_type 'b200', and the host
wrappers `filter_loop` / `smoother_loop` that feed them.

Mirrors (paths relative to /root/reference/src/lib/stgp/computation/filters/):
  kalman_filter.py:487-547  filter_loop(data, prior, R, R_inv, filter_type, train_test_mask, train_index)
  kalman_filter.py:439-485  filter(data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index)
  rts_smoother.py:194-219   smoother_loop(data, model, filter_res, full_state, filter_type)
  rts_smoother.py:162-192   smoother(data, model, filter_res, dt, X_t, X_s, full_state)
Same names, argument meaning and return structure; results are torch CUDA tensors (the analogue of
the reference's device-resident jax arrays).  A leading batch axis B is accepted on Y / R and on the
prior (`BatchedMaternSDE`), and is carried through to the outputs when present.
"""
import warnings

import numpy as np
import torch

from . import ops
from . import settings
from .dispatch import dispatch, evoke
from .sdes import PDE, BatchedMaternSDE


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("physs_gp_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


_PIN_LIMIT = 256 << 20      # larger pageable arrays are copied as they are (their copy time dwarfs the stall)


_SMALL_LIMIT = 1 << 20      # host arrays up to this size are cached on the device by content
_small_cache = {}           # (digest, shape, device) -> (device tensor, copy-done event)
captured_status = []        # device flags of parallel-in-time filters enqueued while a CUDA graph is captured


def _to_dev(x, dev):
    """Host -> device without stalling the host: a copy from PAGEABLE memory synchronises the stream before it
    starts (the host then waits for every kernel already queued on it -- measured 10 ms per call in a
    pipelined batch), so host arrays go through a pinned staging tensor (torch's caching host allocator keeps
    it alive until the copy has run) and the copy is stream-ordered like everything else.

    Small numpy arrays (time axes, H, hyper-parameters: the same few arrays on every call of a training loop) are
    cached on the device keyed by a digest of their CONTENT, so a steady-state CVI step issues no host -> device
    copy at all -- which is also what makes the step capturable in a CUDA graph (VGP.compile_step).  The cached
    tensors are inputs; nothing on the path writes to them."""
    if not isinstance(x, torch.Tensor):
        a = np.ascontiguousarray(x, dtype=np.float64)
        if a.nbytes <= _SMALL_LIMIT:
            import hashlib
            key = (hashlib.blake2b(a.tobytes(), digest_size=16).digest(), a.shape, dev.index)
            hit = _small_cache.get(key)
            if hit is None:
                if len(_small_cache) >= 512:
                    _small_cache.clear()
                t = torch.as_tensor(a).pin_memory().to(device=dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                _small_cache[key] = hit = (t, ev)
            elif not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(dev).wait_event(hit[1])     # another stream may have staged it
            return hit[0]
        x = torch.as_tensor(a)
    if x.device.type == 'cpu' and not x.is_pinned() and x.numel() * 8 <= _PIN_LIMIT:
        x = x.to(torch.float64).pin_memory()
    return x.to(device=dev, dtype=torch.float64, non_blocking=True)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)


def lower_prior(prior, X_s, dts, dev, sequential=True):
    """Turn a prior object into what the kernels consume.

    Returns (discs, m0, P0, H): one `ops.Disc` per dt array in `dts` (filter and smoother use
    different dt conventions), m0 [*, d], P0 [*, d, d] device tensors and H as a numpy [m, d] array.
    """
    if isinstance(prior, BatchedMaternSDE):
        # the prior's device tensors are cached on the prior object (keyed by a fingerprint of its
        # hyper-parameters): a CVI iteration calls the filter + smoother four times with the same prior, and
        # re-staging lam / P_inf through pinned memory on every call was a measurable part of its step time
        key = (str(dev), hash(prior.ls.tobytes()), hash(prior.var.tobytes()))
        cache = getattr(prior, "_b200_cache", None)
        cur = torch.cuda.current_stream(dev)
        if cache is None or cache[0] != key:
            cache = (key, _to_dev(prior.lam(), dev), _to_dev(prior.P_inf(), dev), _to_dev(prior.m_inf(), dev), cur)
            prior._b200_cache = cache
        elif cache[4] != cur and not torch.cuda.is_current_stream_capturing():
            cur.wait_stream(cache[4])                  # the staging copies were enqueued on another stream
            # (a graph capture is preceded by a device-wide sync in VGP.compile_step: nothing to wait for, and a
            #  dependency on uncaptured work would invalidate the capture)
        _, lam, Pinf, m0, _ = cache
        disc = ops.Disc.matern(prior.nblk, lam, Pinf)
        return [disc for _ in dts], m0, Pinf, prior.H()
    P_inf = np.asarray(prior.P_inf(None, X_s, None), np.float64)
    m_inf = np.asarray(prior.m_inf(None, X_s, None), np.float64).reshape(1, -1)
    H = np.asarray(prior.H(None, X_s, None), np.float64)
    d = P_inf.shape[0]
    blocks = prior.ss_blocks() if hasattr(prior, "ss_blocks") else None
    P0 = _to_dev(P_inf[None], dev)
    m0 = _to_dev(m_inf, dev)
    if blocks is not None and d <= 4 and any(np.signbit(l) for _, l in blocks):
        blocks = None   # oscillator blocks (periodic kernels) are evaluated on chip by the lane-group kernels only
    if blocks is not None and len({s for s, _ in blocks}) == 1:
        s = blocks[0][0]
        mask = np.kron(np.eye(len(blocks)), np.ones([s, s]))
        if np.all(P_inf * (1 - mask) == 0.0):
            lam = _to_dev(np.array([[l for _, l in blocks]]), dev)
            disc = ops.Disc.matern(len(blocks), lam, P0)
            if ops.kf_supported(d, 1, disc):
                return [disc for _ in dts], m0, P0, H
    iwp = prior.iwp_blocks() if (sequential and hasattr(prior, "iwp_blocks")) else None
    if iwp is not None and len(iwp) == 1 and 2 <= iwp[0][0] <= 4 and iwp[0][0] == d:
        # one integrated-Wiener block: closed-form A_k, Q_k on chip (PHYSS_DISC_IWP), sequential kernels only
        disc = ops.Disc.iwp(_to_dev(np.array([[iwp[0][1]]]), dev))
        return [disc for _ in dts], m0, P0, H
    # generic route: evaluate the reference prior API once per distinct dt (host), ship A_k, Q_k
    discs = []
    for dt in dts:
        dt_np = _np(dt).reshape(-1)
        uniq, inv = np.unique(dt_np, return_inverse=True)
        A_u = np.stack([np.asarray(prior.expm(X_s, float(u)), np.float64) for u in uniq])
        Q_u = np.stack([np.asarray(prior.Q(float(u), A_u[i], P_inf, X_s), np.float64)
                        for i, u in enumerate(uniq)])
        discs.append(ops.Disc.given(_to_dev(A_u[inv], dev), _to_dev(Q_u[inv], dev)))
    return discs, m0, P0, H


def lower_prior_big(prior, X_s, dt, dev):
    """Large-block route (d > settings.big_block_min_dim, one series): prior.expm / prior.Q are evaluated once
    per DISTINCT step size on the host (the reference API) and shipped as [nA, d, d] with a host index."""
    P_inf = np.asarray(prior.P_inf(None, X_s, None), np.float64)
    m_inf = np.asarray(prior.m_inf(None, X_s, None), np.float64).reshape(-1)
    H = np.asarray(prior.H(None, X_s, None), np.float64)
    dt_np = _np(dt).reshape(-1)
    uniq, inv = np.unique(dt_np, return_inverse=True)
    A_u = np.stack([np.asarray(prior.expm(X_s, float(u)), np.float64) for u in uniq])
    Q_u = np.stack([np.asarray(prior.Q(float(u), A_u[i], P_inf, X_s), np.float64) for i, u in enumerate(uniq)])
    return ops.BigDisc(_to_dev(A_u, dev), _to_dev(Q_u, dev), inv), _to_dev(m_inf, dev), _to_dev(P_inf, dev), H


def _kron_parts(prior, X_s):
    """(temporal kernel, Ks, ds) when `prior` is a plain LTI_SDE over ONE separable spatio-temporal kernel
    (kernels/kernel.py:213-265) observed through H = I (x) [1 0 ..] -- the shape the hand-written large-block
    kernels (physs_kf_filter_kron_f64) cover -- else None."""
    from .kernels import SpatioTemporalSeperableKernel
    from .sdes import LTI_SDE
    if not settings.kron_kernels or type(prior) is not LTI_SDE or len(prior.gp.parent) != 1:
        return None
    k = prior.gp.parent[0].kernel
    if not isinstance(k, SpatioTemporalSeperableKernel):
        return None
    h = np.asarray(k.k1.to_ss()[3], np.float64).reshape(-1)
    ds = h.shape[0]
    if ds > 4 or h[0] != 1.0 or np.any(h[1:] != 0.0):
        return None
    return k.k1, k.Ks, ds


def lower_prior_kron(parts, prior, X_s, dt, dev):
    """Separable route: the TEMPORAL kernel's expm / Q (ds x ds) once per distinct step size; the d x d Kronecker
    products the reference builds every step (kernel.py:247-254) are never formed."""
    k1, Ks, ds = parts
    Pinf_t = np.asarray(k1.to_ss()[5], np.float64)
    dt_np = _np(dt).reshape(-1)
    uniq, inv = np.unique(dt_np, return_inverse=True)
    At = np.stack([np.asarray(k1.expm(float(u)), np.float64) for u in uniq])
    Qt = np.stack([np.asarray(k1.Q(float(u), At[i], Pinf_t), np.float64) for i, u in enumerate(uniq)])
    m_inf = np.asarray(prior.m_inf(None, X_s, None), np.float64).reshape(-1)
    disc = ops.KronDisc(_to_dev(At, dev), _to_dev(Qt, dev), torch.as_tensor(inv.astype(np.int32)), _to_dev(Ks, dev))
    return disc, _to_dev(m_inf, dev), _to_dev(np.kron(Ks, Pinf_t), dev)


def _use_big(prior, X_s, batched_B):
    if isinstance(prior, BatchedMaternSDE) or batched_B != 1:
        return False
    d = np.asarray(prior.P_inf(None, X_s, None)).shape[0]
    return d > settings.big_block_min_dim


def _is_identity(H):
    return H.shape[0] == H.shape[1] and np.array_equal(H, np.eye(H.shape[0]))


def _filter_impl(parallel, data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index):
    """B200 backend of evoke('filter', filter_type) -- kalman_filter.py:439-485.

    Y [T, m, 1] (or [B, T, m, 1]), lik_mat R [T, m, m] (or [B, T, m, m] / broadcastable),
    dt [T] with dt[0] = 0.  Returns (lml, {'m': [T, d, 1], 'P': [T, d, d]}) with a leading B when
    the inputs were batched."""
    dev = _device()
    if not lik_cov_flag:
        # precision-parameterised sites (R_inv): the reference's sequential path raises (kalman_filter.py:67) and
        # only its parallel filter takes them (parallel_kalman_filter.py:34-71,117-141).  Here they are turned into
        # covariances on the device (Cholesky factor-and-solve, as the reference's precision elements do) and the
        # ordinary recursion runs.
        Rinv = _to_dev(lik_mat, dev)
        if Rinv.shape[-1] > 8:
            raise NotImplementedError("precision sites: observation blocks up to 8 x 8")
        lik_mat, lik_cov_flag = ops.spd_inverse(Rinv, 0.0), True
    Yd = _to_dev(Y, dev)
    one = Yd.dim() == 3 or (Yd.dim() == 4 and Yd.shape[0] == 1)
    if one and _use_big(prior, X_s, 1):
        # one series with a large state (BASELINE config 2)
        R = _to_dev(lik_mat, dev)
        lead = Yd.dim() == 4
        if lead:
            Yd = Yd[0]
            R = R[0] if R.dim() == 4 else R
        parts = _kron_parts(prior, X_s)
        if parts is not None:
            # separable prior: hand-written persistent kernels that use the Kronecker structure
            disc, m0, P0 = lower_prior_kron(parts, prior, X_s, _to_dev(dt, dev), dev)
            lml, mf, Pf = ops.kf_filter_kron(Yd[..., 0], R, m0, P0, disc, jitter=settings.jitter)
        else:
            # any other large prior: dense A, Q through the cuBLAS / cuSOLVER-backed library path
            disc, m0, P0, H = lower_prior_big(prior, X_s, _to_dev(dt, dev), dev)
            lml, mf, Pf = ops.kf_filter_big(Yd[..., 0], R, _to_dev(H, dev), m0, P0, disc, jitter=settings.jitter)
        if lead:
            return lml[None], {'m': mf[None, ..., None], 'P': Pf[None]}
        return lml, {'m': mf[..., None], 'P': Pf}
    batched = Yd.dim() == 4 or isinstance(prior, BatchedMaternSDE)
    if Yd.dim() == 3:
        Yd = Yd[None]
    Yd = Yd[..., 0]                                   # [B, T, m]
    if isinstance(prior, BatchedMaternSDE) and Yd.shape[0] == 1 and prior.B > 1:
        Yd = Yd.expand(prior.B, -1, -1)
    if settings.time_major and Yd.shape[0] >= settings.time_major_min_batch and not ops.step_layout(Yd, "Y")[1]:
        # B200 layout: store the batch time-major ([T, B, m] in memory, logical shape unchanged) so that
        # the 32 series of a warp read / write contiguous spans; every output follows Y's memory order
        Yd = Yd.transpose(0, 1).contiguous().transpose(0, 1)
    dtd = _to_dev(dt, dev)
    (disc,), m0, P0, H = lower_prior(prior, X_s, [dtd], dev, sequential=not parallel)
    R = _to_dev(lik_mat, dev)
    Hd = None if _is_identity(H) else _to_dev(H, dev)
    if isinstance(prior, PDE):
        # collocation (EKF) step: kf_predict_step(PDE, 'sequential'), kalman_filter.py:340-427.  Sequential only
        # (the reference has no parallel-in-time form of it either).
        res = prior.residuals
        T = Yd.shape[1]
        forcing = None
        if any(r.forcing is not None for r in res):
            forcing = _to_dev(np.stack([np.zeros(T) if r.forcing is None else r.forcing for r in res]), dev)
        terms = [(p_, k, i, c) for p_, r in enumerate(res) for k, i, c in r.terms]
        bnd = None
        if prior.boundary_conditions is not None:
            bc = np.asarray(prior.boundary_conditions, np.float64)
            if train_index is not None:
                bc = bc[np.asarray(train_index)]                 # kalman_filter.py:464-467
            bnd = _to_dev(bc.reshape(1, T, -1), dev).expand(Yd.shape[0], -1, -1).contiguous()
        lml, mf, Pf = ops.kf_filter_colloc(dtd, Yd, R, Hd, m0, P0, disc, np.stack([r.w for r in res]), terms,
                                           forcing=forcing, y_pseudo=prior.psuedo_observations()[:, 0], boundary=bnd,
                                           observe_data=prior.observe_data, jitter=settings.jitter)
    elif parallel:
        lml, mf, Pf, status = ops.pscan_filter(dtd, Yd, R, Hd, m0, P0, disc, chunk_len=settings.pscan_chunk_len,
                                               jitter=settings.jitter, polish=settings.pscan_polish,
                                               return_status=True)
        if torch.cuda.is_current_stream_capturing():
            captured_status.append(status)          # no host sync inside a graph: the owner of the graph reads it
        elif settings.pscan_check_status and int(status.item()) != 0:
            # Some chunk did not reconcile with the jittered sequential recursion within the fix-up passes (slowly
            # mixing filter / chunks shorter than its memory).  Every further pass contracts the boundary error by
            # the forgetting over one chunk and costs a few steps per converged chunk, so first retry with four
            # times the passes (tens of ms) before giving the result up for the sequential kernels (seconds for a
            # single long series): no cliff for a merely slow-mixing model.
            base = settings.pscan_polish if settings.pscan_polish is not None else 4
            lml, mf, Pf, status = ops.pscan_filter(dtd, Yd, R, Hd, m0, P0, disc, chunk_len=settings.pscan_chunk_len,
                                                   jitter=settings.jitter, polish=4 * max(base, 1),
                                                   return_status=True)
            if int(status.item()) != 0:
                warnings.warn("physs_gp_b200: parallel-in-time filter did not converge to the sequential recursion "
                              "(raise settings.pscan_polish or settings.pscan_chunk_len); using the sequential kernels")
                lml, mf, Pf = ops.kf_filter(dtd, Yd, R, Hd, m0, P0, disc, jitter=settings.jitter)
    else:
        lml, mf, Pf = ops.kf_filter(dtd, Yd, R, Hd, m0, P0, disc, jitter=settings.jitter)
    if batched:
        return lml, {'m': mf[..., None], 'P': Pf}
    return lml[0], {'m': mf[0][..., None], 'P': Pf[0]}


@dispatch('b200')
def filter(data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index):
    """filter_type='b200': sequential-in-time kernels (one thread / lane group per series)."""
    return _filter_impl(False, data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index)


@dispatch('b200_parallel')
def filter(data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index):  # noqa: F811
    """filter_type='b200_parallel': parallel-in-time chunked associative scan (the counterpart of the
    reference's filter('parallel'), parallel_kalman_filter.py:225-336).  Returns the SEQUENTIAL result
    (SURVEY quirk Q1); use it for long series / small batches."""
    return _filter_impl(True, data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index)


def _auto_parallel(B, T):
    """filter_type='b200_auto' (the counterpart of the reference's parallel='auto', zoo/sde_diff.py:370-378):
    go parallel in time when the batch alone cannot fill the GPU (one thread / lane group per series needs
    ~10^4 series) and the series is long enough to cut into chunks."""
    return B < settings.auto_parallel_max_batch and T >= settings.auto_parallel_min_steps


@dispatch('b200_auto')
def filter(data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index):  # noqa: F811
    B = Y.shape[0] if Y.dim() == 4 else (prior.B if isinstance(prior, BatchedMaternSDE) else 1)
    return _filter_impl(_auto_parallel(B, Y.shape[-3]), data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag,
                        train_test_mask, train_index)


def _smoother_impl(parallel, data, model, filter_res, dt, X_t, X_s, full_state):
    """B200 backend of evoke('smoother', filter_type) -- rts_smoother.py:162-192.
    dt [T] with dt[k] = t_{k+1} - t_k, dt[T-1] = 0."""
    dev = _device()
    mf = _to_dev(filter_res['m'], dev)[..., 0]
    Pf = _to_dev(filter_res['P'], dev)
    one = mf.dim() == 2 or (mf.dim() == 3 and mf.shape[0] == 1)
    if one and _use_big(model, X_s, 1):
        lead = mf.dim() == 3
        parts = _kron_parts(model, X_s)
        if parts is not None:
            disc, _, _ = lower_prior_kron(parts, model, X_s, _to_dev(dt, dev), dev)
            ms, Ps = ops.rts_smooth_kron(mf[0] if lead else mf, Pf[0] if lead else Pf, disc,
                                         project=not full_state, jitter=settings.jitter)
        else:
            disc, _, _, H = lower_prior_big(model, X_s, _to_dev(dt, dev), dev)
            Hout = None if (full_state or _is_identity(H)) else _to_dev(H, dev)
            ms, Ps = ops.rts_smooth_big(mf[0] if lead else mf, Pf[0] if lead else Pf, disc, Hout=Hout,
                                        jitter=settings.jitter)
        if lead:
            return ms[None, ..., None], Ps[None]
        return ms[..., None], Ps
    batched = mf.dim() == 3
    if not batched:
        mf, Pf = mf[None], Pf[None]
    dtd = _to_dev(dt, dev)
    (disc,), _, _, H = lower_prior(model, X_s, [dtd], dev, sequential=not parallel)
    Hout = None if (full_state or _is_identity(H)) else _to_dev(H, dev)
    if parallel:
        ms, Ps = ops.pscan_smooth(dtd, mf, Pf, disc, Hout=Hout, chunk_len=settings.pscan_chunk_len,
                                  jitter=settings.jitter)
    else:
        ms, Ps = ops.rts_smooth(dtd, mf, Pf, disc, Hout=Hout, jitter=settings.jitter)
    if batched:
        return ms[..., None], Ps
    return ms[0][..., None], Ps[0]


@dispatch('b200')
def smoother(data, model, filter_res, dt, X_t, X_s, full_state):
    return _smoother_impl(False, data, model, filter_res, dt, X_t, X_s, full_state)


@dispatch('b200_parallel')
def smoother(data, model, filter_res, dt, X_t, X_s, full_state):  # noqa: F811
    """Counterpart of smoother('parallel') (parallel_rts_smoother.py:57-103); honours full_state
    (the reference's parallel smoother ignores it)."""
    return _smoother_impl(True, data, model, filter_res, dt, X_t, X_s, full_state)


def _time_axis(data, dev):
    return _to_dev(data.X_time, dev).reshape(-1)


def filter_loop(data, prior, R=None, R_inv=None, filter_type='b200', train_test_mask=None, train_index=None):
    """kalman_filter.py:487-547: dt = [0, diff(t)], Y_st [Nt, P, Ns] -> [Nt, P*Ns, 1], pick R / R_inv,
    resolve the backend by `filter_type`."""
    dev = _device()
    X_t = _time_axis(data, dev)
    X_s = data.X_space
    dt = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), X_t[1:] - X_t[:-1]])
    Y = _to_dev(data.Y_st, dev)
    Y = Y.reshape(*Y.shape[:-2], -1)[..., None]       # [..., Nt, P*Ns, 1]
    if R_inv is not None:
        lik_cov_flag, lik_mat = False, R_inv
    else:
        lik_cov_flag, lik_mat = True, R
    filter_fn = evoke('filter', filter_type)
    return filter_fn(data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag, train_test_mask, train_index)


def smoother_loop(data, model, filter_res, full_state=False, filter_type='b200'):
    """rts_smoother.py:194-219: dt = [diff(t), 0]."""
    dev = _device()
    X_t = _time_axis(data, dev)
    X_s = data.X_space
    dt = torch.cat([X_t[1:] - X_t[:-1], torch.zeros(1, dtype=torch.float64, device=dev)])
    smoother_fn = evoke('smoother', filter_type)
    return smoother_fn(data, model, filter_res, dt, X_t, X_s, full_state)


def filter_smooth_fused(data, prior, R=None, R_inv=None, filter_type='b200'):
    """`filter_loop` + `smoother_loop(full_state=False)` (sde_gp.py:231-253) as ONE C-ABI call that never
    materialises the filtered moments (`physs_kf_filter_smooth_packed_f64`): the hand-over between the two passes
    is a workspace of packed rows [m | triu(P)].  Returns (lml, mu, var) -- what the two loops return (lml bitwise,
    the projected moments to 1e-14) -- or None when the problem is outside what the packed call covers (sequential register kernels, state dim <= 4,
    a time-major batch, covariance-parameterised noise); the caller then takes the two loops."""
    if not settings.fused_packed or R_inv is not None or R is None or isinstance(prior, PDE):
        return None
    X_s = data.X_space
    d = prior.d if isinstance(prior, BatchedMaternSDE) else np.asarray(prior.P_inf(None, X_s, None)).shape[0]
    if d > 4:
        return None
    dev = _device()
    X_t = _time_axis(data, dev)
    Y = _to_dev(data.Y_st, dev)
    Y = Y.reshape(*Y.shape[:-2], -1)                  # [..., Nt, P*Ns]
    batched = Y.dim() == 3 or isinstance(prior, BatchedMaternSDE)
    if not batched:
        return None                                   # one series: nothing to coalesce over
    if filter_type == 'b200_auto':
        Bq = Y.shape[0] if Y.dim() == 3 else prior.B
        if _auto_parallel(Bq, Y.shape[-2]):
            return None
    elif filter_type != 'b200':
        return None
    if Y.dim() == 2:
        Y = Y[None]
    if isinstance(prior, BatchedMaternSDE) and Y.shape[0] == 1 and prior.B > 1:
        Y = Y.expand(prior.B, -1, -1)
    if not (settings.time_major and Y.shape[0] >= settings.time_major_min_batch):
        return None
    if not ops.step_layout(Y, "Y")[1]:
        Y = Y.transpose(0, 1).contiguous().transpose(0, 1)
    zero = torch.zeros(1, dtype=torch.float64, device=dev)
    diff = X_t[1:] - X_t[:-1]
    dt_f, dt_s = torch.cat([zero, diff]), torch.cat([diff, zero])
    (disc_f, disc_s), m0, P0, H = lower_prior(prior, X_s, [dt_f, dt_s], dev, sequential=True)
    if not ops.kf_filter_smooth_packed_supported(Y, d, Y.shape[-1], disc_f):
        return None
    ident = _is_identity(H)
    Hd = None if ident else _to_dev(H, dev)
    lml, ms, Ps = ops.kf_filter_smooth_packed(dt_f, dt_s, Y, _to_dev(R, dev), Hd, m0, P0, disc_f, disc_s, Hout=Hd,
                                              jitter=settings.jitter)
    return lml, ms[..., None], Ps


@dispatch('b200_auto')
def smoother(data, model, filter_res, dt, X_t, X_s, full_state):  # noqa: F811
    mf = filter_res['m']
    B, T = (mf.shape[0], mf.shape[1]) if mf.dim() == 4 else (1, mf.shape[0])
    return _smoother_impl(_auto_parallel(B, T), data, model, filter_res, dt, X_t, X_s, full_state)
