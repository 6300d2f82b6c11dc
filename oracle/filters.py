"""Oracle: Kalman filter / RTS smoother, sequential and parallel-in-time (test infrastructure).

Follows (paths relative to /root/reference/src/lib/stgp/computation/filters/):
  kalman_filter.py:144-211    kf_update_step (masking, jittered gain solve, P - K S K^T, lml)
  kalman_filter.py:214-241    kf_predict_step(LTI_SDE)
  kalman_filter.py:340-427    kf_predict_step(PDE): collocation (EKF) step -- filter_pde_sequential
  kalman_filter.py:439-485    filter('sequential'): scan from (m_inf, P_inf), lml = sum
  kalman_filter.py:487-547    filter_loop: dt = [0, diff(t)], Y -> [T, m, 1]
  rts_smoother.py:48-65       rts_smoother_step (jittered chol of P_pred)
  rts_smoother.py:69-106      rts_step_wrapper(LTI_SDE)
  rts_smoother.py:162-192     smoother('sequential'): reverse scan, last step prepended, H projection
  rts_smoother.py:194-219     smoother_loop: dt = [diff(t), 0]
  parallel_kalman_filter.py:73-100,143-175,178-220,225-336   parallel filter (elements, operator, lml)
  parallel_rts_smoother.py:21-103                            parallel smoother

Reference quirks reproduced on purpose are listed in SURVEY.md section 8 (Q1-Q4).
"""
import numpy as np

from . import linalg as la


# ------------------------------------------------------------------ sequential filter

def kf_update_step(m_, P_, H, R, y, jitter=la.JITTER, innovation=None):
    """kalman_filter.py:144-211.  y: [m,1] with NaN = missing.  Returns m, P, lml_k.
    `innovation` (the reference's last argument): the predicted observation; H m_ for linear observations, the
    non-linear residual g(m_) for the collocation update of the PDE filter (:413)."""
    mask = (~np.isnan(y)).astype(int)              # nan_utils.py:13-20
    y0 = np.nan_to_num(y)
    M = np.tile(mask, [1, y.shape[0]]) * np.eye(y.shape[0])
    if innovation is None:
        innovation = H @ m_
    mu = M @ innovation
    var = M @ H @ P_ @ H.T @ M.T
    v = y0 - mu
    S = var + R                                     # R is NOT masked
    K = la.solve(S, M @ H @ P_, jitter).T           # jitter only inside this solve
    m = m_ + K @ v
    P = P_ - K @ S @ K.T                            # un-jittered S, not Joseph form
    lml = la.log_gaussian_with_mask(y0, mu, S, mask[:, 0])
    return m, P, lml


def filter_sequential(prior, X_time, Y, R, jitter=la.JITTER):
    """filter_loop + filter('sequential').

    X_time [T]; Y [T, m] (NaN = missing); R [T, m, m].
    Returns lml (scalar), m [T, d, 1], P [T, d, d], lml_k [T]."""
    T = X_time.shape[0]
    dt = np.hstack([np.zeros(1), np.diff(X_time)])              # kalman_filter.py:515
    Ycol = np.reshape(Y, [T, -1])[..., None]                    # :523-528
    m = prior.m_inf()
    P = prior.P_inf()
    P_inf = prior.P_inf()
    H = prior.H()
    ms, Ps, lmls = [], [], []
    for k in range(T):
        A = prior.expm(dt[k])                                   # :230
        Q = prior.Q(dt[k], A, P_inf)                            # :231
        m_ = A @ m
        P_ = A @ P @ A.T + Q
        m, P, l = kf_update_step(m_, P_, H, R[k], Ycol[k], jitter)
        ms.append(m), Ps.append(P), lmls.append(l)
    lmls = np.array(lmls)
    return float(np.sum(lmls)), np.array(ms), np.array(Ps), lmls


# ------------------------------------------------------------------ collocation (EKF) filter

class PointResidual:
    """Point-wise collocation residual of an ODE / PDE on the derivative-augmented state x [d]:

        g(x, k) = w . x + sum_q coef_q * phi_q(x[idx_q]) + forcing[k],     phi in {sin, cos, square, cube}

    with its Jacobian dg/dx = w + sum_q coef_q phi_q'(x[idx_q]) e_idx_q -- what `PDE.forward_g` / `PDE.H_jac`
    (transforms/pdes.py:236-245) evaluate with jax.jacfwd for the point-wise residuals the reference ships
    (Pendulum1D :482-528, DampedPendulum1D :530-597, SimpleODE :424-480, Allen-Cahn's u^3 - u :700-811)."""
    KINDS = {"sin": 0, "cos": 1, "square": 2, "cube": 3, "prod": 4}

    def __init__(self, w, terms=(), forcing=None):
        """terms: (kind, state index, coefficient); kind "prod" is the bilinear coef * x[i] * x[j] of the reference's
        ODE systems (LotkaVolterra, LorenzSystem: transforms/pdes.py:818-1008) with the index i | (j << 8)."""
        self.w = np.asarray(w, float)
        self.terms = [(k, int(i), float(c)) for k, i, c in terms]
        self.forcing = None if forcing is None else np.asarray(forcing, float)

    def g(self, x, k):
        x = np.ravel(x)
        v = float(self.w @ x)
        for kind, i, c in self.terms:
            if kind == "prod":
                v += c * x[i & 255] * x[i >> 8]
                continue
            v += c * {"sin": np.sin, "cos": np.cos, "square": lambda z: z * z, "cube": lambda z: z ** 3}[kind](x[i])
        return v + (0.0 if self.forcing is None else self.forcing[k])

    def jac(self, x):
        x = np.ravel(x)
        J = self.w.copy()
        for kind, i, c in self.terms:
            if kind == "prod":
                J[i & 255] += c * x[i >> 8]
                J[i >> 8] += c * x[i & 255]
                continue
            J[i] += c * {"sin": np.cos, "cos": lambda z: -np.sin(z), "square": lambda z: 2 * z,
                         "cube": lambda z: 3 * z * z}[kind](x[i])
        return J[None, :]


def filter_pde_sequential(prior, residuals, X_time, Y, R, boundary=None, y_pseudo=None, observe_data=True,
                          jitter=la.JITTER):
    """kf_predict_step(PDE, 'sequential') (kalman_filter.py:340-427) inside filter('sequential') (:439-485):
    LTI predict, optional boundary update with R * 0 (:382-391), pseudo-observation update with H = dg/dx at the
    PREDICTED mean, zero noise and innovation g(m_) (:395-414; both evaluated before the boundary update, :378-379),
    then the data update (:417-421).  The step's lml is that of the LAST update executed (the `ys` returned).

    residuals: list of Pc PointResidual (the reference needs Pc >= 2: `np.squeeze(f)[..., None]`, :413, is rank 1
    for a single output).  Y [T, m] data (NaN = missing); boundary [T, m] or None (NaN = no boundary observation
    at that step); y_pseudo [Pc]: the pseudo observations (0, or NaN for "no collocation on this output").
    Returns lml, m [T,d,1], P [T,d,d], lml_k."""
    Pc = len(residuals)
    y_ps = np.zeros(Pc) if y_pseudo is None else np.asarray(y_pseudo, float)
    T = X_time.shape[0]
    dt = np.hstack([np.zeros(1), np.diff(X_time)])
    Ycol = np.reshape(Y, [T, -1])[..., None]
    m, P = prior.m_inf(), prior.P_inf()
    P_inf, H = prior.P_inf(), prior.H()
    ms, Ps, lmls = [], [], []
    for k in range(T):
        A = prior.expm(dt[k])
        Q = prior.Q(dt[k], A, P_inf)
        m_ = A @ m
        P_ = A @ P @ A.T + Q
        f = np.array([[r.g(m_, k)] for r in residuals])
        Hj = np.vstack([r.jac(m_) for r in residuals])
        l = 0.0
        if boundary is not None:
            m_, P_, l = kf_update_step(m_, P_, H, R[k] * 0.0, np.reshape(boundary[k], [-1, 1]), jitter,
                                       innovation=H @ m_)
        m_, P_, l = kf_update_step(m_, P_, Hj, np.zeros((Pc, Pc)), y_ps[:, None], jitter, innovation=f)
        if observe_data:
            m_, P_, l = kf_update_step(m_, P_, H, R[k], Ycol[k], jitter, innovation=H @ m_)
        m, P = m_, P_
        ms.append(m), Ps.append(P), lmls.append(l)
    lmls = np.array(lmls)
    return float(np.sum(lmls)), np.array(ms), np.array(Ps), lmls


# ------------------------------------------------------------------ sequential smoother

def rts_smoother_step(m_f, P_f, m_next, P_next, m_pred, P_pred, A, jitter=la.JITTER):
    """rts_smoother.py:48-65."""
    Lp = la.cholesky(la.add_jitter(P_pred, jitter))
    G = la.cholesky_solve(Lp, A @ P_f).T
    m = m_f + G @ (m_next - m_pred)
    P = P_f + G @ (P_next - P_pred) @ G.T
    return m, P


def smoother_sequential(prior, X_time, m_f, P_f, full_state=False, jitter=la.JITTER):
    """smoother_loop + smoother('sequential').  Returns (H m_s [T,m',1], H P_s H^T [T,m',m'])."""
    T = X_time.shape[0]
    dt = np.hstack([np.diff(X_time), np.zeros(1)])              # rts_smoother.py:209-211
    P_inf = prior.P_inf()
    d = m_f.shape[1]
    Hk = np.eye(d) if full_state else prior.H()                 # :26-35
    m, P = m_f[-1], P_f[-1]
    out_m = [Hk @ m]
    out_P = [Hk @ P @ Hk.T]
    for k in range(T - 2, -1, -1):
        A = prior.expm(dt[k])
        Q = prior.Q(dt[k], A, P_inf)
        m_pred = A @ m_f[k]
        P_pred = A @ P_f[k] @ A.T + Q
        m, P = rts_smoother_step(m_f[k], P_f[k], m, P, m_pred, P_pred, A, jitter)
        out_m.append(Hk @ m)
        out_P.append(Hk @ P @ Hk.T)
    return np.array(out_m[::-1]), np.array(out_P[::-1])


def filter_and_smooth(prior, X_time, Y, R, full_state=False, jitter=la.JITTER):
    """models/sde_gp.py:231-253."""
    lml, m_f, P_f, _ = filter_sequential(prior, X_time, Y, R, jitter)
    m_s, P_s = smoother_sequential(prior, X_time, m_f, P_f, full_state, jitter)
    return lml, m_s, P_s


# ------------------------------------------------------------------ parallel-in-time forms

def first_filtering_element(m, P, F, Q, H, R, y, jitter=la.JITTER):
    """parallel_kalman_filter.py:73-100."""
    m_ = F @ m
    P_ = F @ P @ F.T + Q
    S1 = H @ P_ @ H.T + R
    K = la.solve(S1, H @ P_.T, jitter).T
    A = np.zeros_like(F)
    b = m_ + K @ (y - H @ m_)
    C = P_ - K @ S1 @ K.T
    S = H @ Q @ H.T + R
    FH_S_inv = la.solve(S, H @ F, jitter).T
    eta = FH_S_inv @ y
    J = FH_S_inv @ H @ F
    return A, b, la.force_symmetric(C), la.force_symmetric(J), eta


def generic_filtering_element(F, Q, H, R, y, jitter=la.JITTER):
    """parallel_kalman_filter.py:143-162."""
    I = np.eye(F.shape[0])
    S = H @ Q @ H.T + R
    K = la.solve(S, H @ Q.T, jitter).T
    A = (I - K @ H) @ F
    b = K @ y
    C = (I - K @ H) @ Q
    eta = F.T @ H.T @ la.solve(S, y, jitter)
    J = F.T @ H.T @ la.solve(S, H @ F, jitter)
    return A, b, C, J, eta


def generic_filtering_element_nan(F, Q):
    """parallel_kalman_filter.py:164-172."""
    d = F.shape[0]
    return F, np.zeros([d, 1]), Q, np.zeros_like(Q), np.zeros([d, 1])


def filtering_operator(x1, x2):
    """parallel_kalman_filter.py:178-220 (parallel_kf_force_linear_solve = False branch)."""
    A_i, b_i, C_i, J_i, eta_i = x1
    A_j, b_j, C_j, J_j, eta_j = x2
    I = np.eye(A_i.shape[1])
    Aj_tmp = np.linalg.solve((I + C_i @ J_j).T, A_j.T).T
    A = Aj_tmp @ A_i
    C = Aj_tmp @ C_i @ A_j.T + C_j
    b = Aj_tmp @ (b_i + C_i @ eta_j) + b_j
    Ai_tmp = np.linalg.solve((I + J_j @ C_i).T, A_i).T
    eta = Ai_tmp @ (eta_j - J_j @ b_i) + eta_i
    J = Ai_tmp @ J_j @ A_i + J_i
    return A, b, la.force_symmetric(C), la.force_symmetric(J), eta


def filter_parallel_reference(prior, X_time, Y, R, jitter=la.JITTER):
    """filter('parallel') bug-for-bug (parallel_kalman_filter.py:225-336), incl. quirks Q1/Q2:
    element 0 uses Q = P_inf on top of A_0 P_inf A_0^T, masks are whole-step only, and the lml
    mask is computed after nan_to_num (so it is all-ones).  The associative scan is evaluated
    as a left fold, which equals jax.lax.associative_scan up to floating-point re-association."""
    T = X_time.shape[0]
    dt = np.hstack([np.zeros(1), np.diff(X_time)])
    Ycol = np.reshape(Y, [T, -1])[..., None]
    P_inf, m_inf, H = prior.P_inf(), prior.m_inf(), prior.H()
    A_arr = [prior.expm(dt[k]) for k in range(T)]
    Q_arr = [prior.Q(dt[k], A_arr[k], P_inf) for k in range(T)]
    mask = np.array([int(np.any(~np.isnan(Ycol[k]))) for k in range(T)])
    Y0 = np.nan_to_num(Ycol)
    elems = []
    for k in range(T):
        if k == 0:
            if mask[0]:
                e = first_filtering_element(m_inf, P_inf, A_arr[0], P_inf, H, R[0], Y0[0], jitter)
            else:
                d = m_inf.shape[0]
                e = (np.zeros([d, d]), m_inf, P_inf, np.zeros([d, d]), np.zeros([d, 1]))
        elif mask[k]:
            e = generic_filtering_element(A_arr[k], Q_arr[k], H, R[k], Y0[k], jitter)
        else:
            e = generic_filtering_element_nan(A_arr[k], Q_arr[k])
        elems.append(e)
    res = [elems[0]]
    for k in range(1, T):
        res.append(filtering_operator(res[-1], elems[k]))
    m_f = np.array([r[1] for r in res])
    P_f = np.array([r[2] for r in res])
    fm = [m_inf] + [r[1] for r in res[:-1]]
    fP = [P_inf] + [r[2] for r in res[:-1]]
    lml = 0.0
    for k in range(T):
        mu = H @ A_arr[k] @ fm[k]
        S = H @ A_arr[k] @ fP[k] @ A_arr[k].T @ H.T + H @ Q_arr[k] @ H.T
        ones = np.ones(Y0[k].shape[0], dtype=int)      # mask of an already nan_to_num'ed Y
        lml += la.log_gaussian_with_mask(Y0[k], mu, S + R[k], ones)
    return lml, m_f, P_f


def generic_smoothing_element(F, Q, m, P, jitter=la.JITTER):
    """parallel_rts_smoother.py:25-37."""
    Pp = F @ P @ F.T + Q
    Lp = la.cholesky(la.add_jitter(Pp, jitter))
    E = la.cholesky_solve(Lp, F @ P).T
    g = m - E @ F @ m
    L = P - E @ Pp @ E.T
    return E, g, la.force_symmetric(L)


def smoothing_operator(x1, x2):
    """parallel_rts_smoother.py:39-55 (arguments arrive reversed)."""
    E_i, g_i, L_i = x2
    E_j, g_j, L_j = x1
    E = E_i @ E_j
    g = E_i @ g_j + g_i
    L = E_i @ L_j @ E_i.T + L_i
    return E, g, la.force_symmetric(L)


def smoother_parallel_reference(prior, X_time, m_f, P_f, jitter=la.JITTER):
    """smoother('parallel') (parallel_rts_smoother.py:57-103); always projects with H (quirk Q3)."""
    T = X_time.shape[0]
    dt = np.hstack([np.diff(X_time), np.zeros(1)])
    P_inf, H = prior.P_inf(), prior.H()
    elems = []
    for k in range(T - 1):
        A = prior.expm(dt[k])
        Q = prior.Q(dt[k], A, P_inf)
        elems.append(generic_smoothing_element(A, Q, m_f[k], P_f[k], jitter))
    elems.append((np.zeros_like(P_f[-1]), m_f[-1], P_f[-1]))
    res = [None] * T
    res[-1] = elems[-1]
    for k in range(T - 2, -1, -1):
        res[k] = smoothing_operator(res[k + 1], elems[k])
    m = np.array([H @ r[1] for r in res])
    P = np.array([H @ r[2] @ H.T for r in res])
    return m, P
