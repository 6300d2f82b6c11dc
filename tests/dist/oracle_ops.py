"""CPU stand-in for physs_gp_b200.ops' time-shard building blocks, built ONLY from the numpy oracle
(oracle/filters.py elements and associative operators).  Test infrastructure: lets the collective plumbing
of physs_gp_b200/timeshard.py run under gloo on CPU (world_size 2) without a GPU."""
import numpy as np
import torch

from oracle import filters as of


class Disc:
    """Per-step transitions of the local range, given explicitly: A, Q [T, d, d]."""
    def __init__(self, A, Q):
        self.A, self.Q = A, Q


def _np(x):
    return None if x is None else x.detach().cpu().numpy()


def _t(x):
    return torch.as_tensor(np.ascontiguousarray(x))


def pscan_workspace(B, T, d, chunk_len, dev):
    return {}


def _elements(Y, R, H, disc, jitter):
    T = Y.shape[0]
    els = []
    for k in range(T):
        y = Y[k][:, None]
        if np.all(np.isnan(y)):
            els.append(of.generic_filtering_element_nan(disc.A[k], disc.Q[k]))
        else:
            els.append(of.generic_filtering_element(disc.A[k], disc.Q[k], H, R[k], np.nan_to_num(y), jitter))
    return els


def pscan_filter_local(dt, Y, R, H, m0, P0, disc, chunk_len, ws, jitter=0.0, stream=None, out=None):
    Yn, Rn, Hn = _np(Y), _np(R), _np(H)
    B, d = Yn.shape[0], disc.A.shape[-1]
    totals = []
    for b in range(B):
        els = _elements(Yn[b], Rn[b], Hn[0], disc, jitter)
        acc = els[0]
        for e in els[1:]:
            acc = of.filtering_operator(acc, e)
        A, bb, C, J, eta = acc
        totals.append(np.concatenate([A.ravel(), C.ravel(), J.ravel(), bb.ravel(), eta.ravel()]))
    ws['filter'] = True
    res = _t(np.stack(totals))
    if out is not None:
        out.copy_(res)                  # the product's kernel writes the gather slot in place; mirror that
        return out
    return res


def _unpack_filter(e, d):
    dd = d * d
    return (e[:dd].reshape(d, d), e[3 * dd:3 * dd + d].reshape(d, 1), e[dd:2 * dd].reshape(d, d),
            e[2 * dd:3 * dd].reshape(d, d), e[3 * dd + d:].reshape(d, 1))


def pscan_filter_fold(totals, m0, P0, stream=None):
    tn, mn, Pn = _np(totals), _np(m0), _np(P0)
    K, B = tn.shape[0], tn.shape[1]
    d = Pn.shape[-1]
    mo, Po = np.zeros((B, d)), np.zeros((B, d, d))
    for b in range(B):
        st = (np.zeros((d, d)), mn[b].reshape(d, 1), Pn[b], np.zeros((d, d)), np.zeros((d, 1)))
        for k in range(K):
            st = of.filtering_operator(st, _unpack_filter(tn[k, b], d))
        mo[b], Po[b] = st[1][:, 0], st[2]
    return _t(mo), _t(Po)


def pscan_filter_finish(dt, Y, R, H, m0, P0, disc, chunk_len, ws, start=None, jitter=0.0, polish=None, out=None,
                        stream=None, **kw):
    assert ws.get('filter'), "finish without local"
    Yn, Rn, Hn = _np(Y), _np(R), _np(H)
    B, T, d = Yn.shape[0], Yn.shape[1], disc.A.shape[-1]
    mf, Pf, lml = np.zeros((B, T, d)), np.zeros((B, T, d, d)), np.zeros(B)
    for b in range(B):
        m = (_np(start[0])[b] if start is not None else np.broadcast_to(_np(m0), (B, d))[b]).reshape(d, 1)
        P = _np(start[1])[b] if start is not None else np.broadcast_to(_np(P0), (B, d, d))[b]
        for k in range(T):
            m_ = disc.A[k] @ m
            P_ = disc.A[k] @ P @ disc.A[k].T + disc.Q[k]
            m, P, l = of.kf_update_step(m_, P_, Hn[0], Rn[b, k], Yn[b, k][:, None], jitter)
            mf[b, k], Pf[b, k] = m[:, 0], P
            lml[b] += l
    return _t(lml), _t(mf), _t(Pf), torch.zeros(1, dtype=torch.int32)


def pscan_smooth_local(dt, mf, Pf, disc, chunk_len, ws, jitter=0.0, stream=None):
    mn, Pn = _np(mf), _np(Pf)
    B, T, d = mn.shape
    totals = []
    for b in range(B):
        acc = None
        for k in range(T - 1, -1, -1):
            e = of.generic_smoothing_element(disc.A[k], disc.Q[k], mn[b, k][:, None], Pn[b, k], jitter)
            acc = e if acc is None else of.smoothing_operator(acc, e)
        E, g, L = acc
        totals.append(np.concatenate([E.ravel(), L.ravel(), g.ravel()]))
    ws['smooth'] = True
    return _t(np.stack(totals))


def pscan_smooth_fold(totals, m_end, P_end, stream=None):
    tn, mn, Pn = _np(totals), _np(m_end), _np(P_end)
    K, B = tn.shape[0], tn.shape[1]
    d = Pn.shape[-1]
    dd = d * d
    mo, Po = np.zeros((B, d)), np.zeros((B, d, d))
    for b in range(B):
        st = (np.zeros((d, d)), mn[b].reshape(d, 1), Pn[b])
        for k in range(K - 1, -1, -1):
            e = (tn[k, b, :dd].reshape(d, d), tn[k, b, 2 * dd:].reshape(d, 1), tn[k, b, dd:2 * dd].reshape(d, d))
            st = of.smoothing_operator(st, e)
        mo[b], Po[b] = st[1][:, 0], st[2]
    return _t(mo), _t(Po)


def pscan_smooth_finish(dt, mf, Pf, disc, chunk_len, ws, start=None, Hout=None, jitter=0.0, out=None, stream=None):
    assert ws.get('smooth'), "finish without local"
    mn, Pn = _np(mf), _np(Pf)
    B, T, d = mn.shape
    ms, Ps = np.zeros((B, T, d)), np.zeros((B, T, d, d))
    for b in range(B):
        if start is None:
            m, P = mn[b, T - 1][:, None], Pn[b, T - 1]
        else:
            m, P = _np(start[0])[b][:, None], _np(start[1])[b]
        for k in range(T - 1, -1, -1):
            A, Q = disc.A[k], disc.Q[k]
            m_pred = A @ mn[b, k][:, None]
            P_pred = A @ Pn[b, k] @ A.T + Q
            m, P = of.rts_smoother_step(mn[b, k][:, None], Pn[b, k], m, P, m_pred, P_pred, A, jitter)
            ms[b, k], Ps[b, k] = m[:, 0], P
    return _t(ms), _t(Ps)
