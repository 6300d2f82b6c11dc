"""Host-side mirror of the reference's Markov-kernel prior API (the part the filter consumes):
`to_ss()`, `expm(dt)`, `Q(dt, A_k, P_inf)`, `state_space_dim()` -- stgp/kernels/kernel.py:200-209,
stgp/kernels/matern.py, stgp/kernels/ss_utils.py.  Same names and argument meaning.

These objects are tiny and T-independent; they run on the host in numpy.  What is new is
`ss_blocks()`: the description the CUDA kernels need to evaluate A_k = expm(F dt_k) in closed form
on chip (block sizes and lam = sqrt(2 nu)/lengthscale per block), so that A_k/Q_k are never
materialised in HBM.
"""
import numpy as np


def _block_diag(mats):
    n = sum(a.shape[0] for a in mats)
    m = sum(a.shape[1] for a in mats)
    out = np.zeros([n, m])
    i = j = 0
    for a in mats:
        out[i:i + a.shape[0], j:j + a.shape[1]] = a
        i += a.shape[0]
        j += a.shape[1]
    return out


class MarkovKernel:
    """stgp/kernels/kernel.py:200-209."""
    _nu2 = None          # 2 nu
    _state_space_dim = None

    def __init__(self, lengthscales=1.0, variance=1.0, input_dim=1):
        ls = np.atleast_1d(np.asarray(lengthscales, dtype=np.float64))
        self.lengthscales = ls
        self.variance = float(variance)
        self.input_dim = input_dim

    def state_space_dim(self):
        return self._state_space_dim

    state_size = state_space_dim

    @property
    def lam(self):
        return np.sqrt(self._nu2) / self.lengthscales[0]

    def Q(self, dt, A_k, P_inf, X_spatial=None):
        return P_inf - A_k @ P_inf @ A_k.T      # kernel.py:207-209

    def P_inf(self):
        return self.to_ss()[5]

    def ss_blocks(self):
        """[(block size, lam)] for the closed-form on-chip discretisation, or None."""
        return [(self._state_space_dim, float(self.lam))]


class Matern12(MarkovKernel):
    """Ornstein-Uhlenbeck; the reference leaves its state-space form unimplemented
    (matern.py:92-107) -- provided here because a size-1 block costs nothing."""
    _nu2 = 1.0
    _state_space_dim = 1

    def to_ss(self, X_spatial=None):
        lam = self.lam
        v = self.variance
        return (np.array([[-lam]]), np.array([[1.0]]), np.array([[2.0 * lam * v]]),
                np.array([[1.0]]), np.zeros([1, 1]), np.array([[v]]))

    def expm(self, dt, X_spatial=None):
        return np.array([[np.exp(-self.lam * dt)]])

    def K(self, X1, X2):
        r = np.abs(X1.reshape(-1, 1) - X2.reshape(1, -1))
        return self.variance * np.exp(-self.lam * r)


class Matern32(MarkovKernel):
    """stgp/kernels/matern.py:53-90 (+ ScaledMatern32 :13-50 when variance != 1);
    ss_utils.py:6-38."""
    _nu2 = 3.0
    _state_space_dim = 2

    def to_ss(self, X_spatial=None):
        lam, v = self.lam, self.variance
        F = np.array([[0.0, 1.0], [-lam * lam, -2.0 * lam]])
        L = np.array([[0.0], [1.0]])
        Qc = np.array([[4.0 * lam ** 3 * v]])
        H = np.array([[1.0, 0.0]])
        return F, L, Qc, H, np.zeros([2, 1]), np.diag([v, lam * lam * v])

    def expm(self, dt, X_spatial=None):
        lam = self.lam
        return np.exp(-dt * lam) * (np.eye(2) + dt * np.array([[lam, 1.0], [-lam * lam, -lam]]))

    def K(self, X1, X2):
        r = self.lam * np.abs(X1.reshape(-1, 1) - X2.reshape(1, -1))
        return self.variance * (1.0 + r) * np.exp(-r)


class Matern52(MarkovKernel):
    """stgp/kernels/matern.py:109-188 (+ ScaledMatern52 :191-266)."""
    _nu2 = 5.0
    _state_space_dim = 3

    def to_ss(self, X_spatial=None):
        lam, v = self.lam, self.variance
        F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3.0 * lam ** 2, -3.0 * lam]])
        L = np.array([[0.0], [0.0], [1.0]])
        Qc = np.array([[16.0 / 3.0 * lam ** 5 * v]])
        H = np.array([[1.0, 0.0, 0.0]])
        kappa = lam * lam * v / 3.0
        Pinf = np.array([[v, 0.0, -kappa], [0.0, kappa, 0.0], [-kappa, 0.0, lam ** 4 * v]])
        return F, L, Qc, H, np.zeros([3, 1]), Pinf

    def expm(self, dt, X_spatial=None):
        lam = self.lam
        x = dt * lam
        M = np.array([
            [lam * (0.5 * x + 1.0), x + 1.0, 0.5 * dt],
            [-0.5 * x * lam ** 2, lam * (1.0 - x), 1.0 - 0.5 * x],
            [lam ** 3 * (0.5 * x - 1.0), lam ** 2 * (x - 3.0), lam * (0.5 * x - 2.0)]])
        return np.exp(-x) * (np.eye(3) + dt * M)

    def K(self, X1, X2):
        r = self.lam * np.abs(X1.reshape(-1, 1) - X2.reshape(1, -1))
        return self.variance * (1.0 + r + r * r / 3.0) * np.exp(-r)


class Matern72(MarkovKernel):
    """stgp/kernels/matern.py:269-341 (ScaledMatern72)."""
    _nu2 = 7.0
    _state_space_dim = 4

    def to_ss(self, X_spatial=None):
        lam, v = self.lam, self.variance
        F = np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0],
                      [-lam ** 4, -4.0 * lam ** 3, -6.0 * lam ** 2, -4.0 * lam]])
        L = np.array([[0.0], [0.0], [0.0], [1.0]])
        Qc = np.array([[32.0 / 5.0 * lam ** 7 * v]])
        H = np.array([[1.0, 0.0, 0.0, 0.0]])
        k1 = lam ** 2 * v / 5.0
        k2 = lam ** 4 * v / 5.0
        Pinf = np.array([[v, 0.0, -k1, 0.0], [0.0, k1, 0.0, -k2],
                         [-k1, 0.0, k2, 0.0], [0.0, -k2, 0.0, lam ** 6 * v]])
        return F, L, Qc, H, np.zeros([4, 1]), Pinf

    def expm(self, dt, X_spatial=None):
        lam = self.lam
        x = dt * lam
        x2 = x * x
        l2, l3 = lam ** 2, lam ** 3
        M = np.array([
            [lam * (1.0 + 0.5 * x + x2 / 6.0), 1.0 + x + 0.5 * x2, 0.5 * dt * (1.0 + x), dt * dt / 6.0],
            [-x2 * l2 / 6.0, lam * (1.0 + 0.5 * x - 0.5 * x2), 1.0 + x - 0.5 * x2, dt * (0.5 - x / 6.0)],
            [l3 * x * (x / 6.0 - 0.5), x * l2 * (0.5 * x - 2.0), lam * (1.0 - 2.5 * x + 0.5 * x2),
             1.0 - x + x2 / 6.0],
            [l2 * l2 * (x - 1.0 - x2 / 6.0), l3 * (3.5 * x - 4.0 - 0.5 * x2),
             l2 * (4.0 * x - 6.0 - 0.5 * x2), lam * (1.5 * x - 3.0 - x2 / 6.0)]])
        return np.exp(-x) * (np.eye(4) + dt * M)

    def K(self, X1, X2):
        r = self.lam * np.abs(X1.reshape(-1, 1) - X2.reshape(1, -1))
        return self.variance * (1.0 + r + 0.4 * r * r + r ** 3 / 15.0) * np.exp(-r)


ScaledMatern32 = Matern32
ScaledMatern52 = Matern52
ScaledMatern72 = Matern72


class WienerVelocity(MarkovKernel):
    """Integrated Wiener process of order q (stgp/kernels/wiener.py:60-149): state (f, f', ..., f^(q)), F the
    shift matrix, spectral density `variance`.  Not stationary: `Q` is its own closed form (:125-149), `P_inf`
    is `stable_state_covariance * I` and only serves as the initial covariance (:97-103)."""

    def __init__(self, q=1, variance=1.0, stable_state_covariance=0.0, m_init=None):
        self.q = int(q)
        self.variance = float(variance)
        self.stable_state_covariance = float(stable_state_covariance)
        self._state_space_dim = self.q + 1
        self.m_init = (np.zeros([self.q + 1, 1]) if m_init is None
                       else np.reshape(np.asarray(m_init, np.float64), [self.q + 1, 1]))
        self.input_dim = 1

    def to_ss(self, X_spatial=None):
        dim, q = self.q + 1, self.q
        F = np.eye(dim, k=1)
        L = np.hstack([np.zeros(dim - 1), [1.0]])[:, None]
        H = np.hstack([[1.0], np.zeros(q)])[None, :]
        return F, L, np.array([[self.variance]]), H, self.m_init, np.eye(dim) * self.stable_state_covariance

    def expm(self, dt, X_spatial=None):
        import math
        dim = self.q + 1
        return np.array([[dt ** (j - i) / math.factorial(j - i) if j >= i else 0.0 for j in range(dim)]
                         for i in range(dim)])

    def Q(self, dt, A_k=None, P_inf=None, X_spatial=None):
        import math
        dim, q = self.q + 1, self.q
        return self.variance * np.array(
            [[dt ** (2 * q + 1 - i - j) / ((2 * q + 1 - i - j) * math.factorial(q - i) * math.factorial(q - j))
              for j in range(dim)] for i in range(dim)])

    def ss_blocks(self):
        return None                                   # not a Matern block

    def iwp_blocks(self):
        """[(block size, spectral density)] for the on-chip integrated-Wiener discretisation (PHYSS_DISC_IWP)."""
        return [(self.q + 1, self.variance)]


IntegratedWiener = WienerVelocity


class ApproxSDEPeriodic_BN(MarkovKernel):
    """stgp/kernels/periodic.py:171-253: periodic covariance as a stack of n_terms + 1 harmonic oscillators,
    d = 2 (n_terms + 1).  Same constructor arguments and `to_ss` / `expm` / `K` as the reference.  The reference
    evaluates `expm(F dt)` with the generic Pade routine per step; F is block-diagonal with blocks j w [[0, -1], [1, 0]],
    so A_k is a stack of plane rotations by j w dt -- which is what `expm` returns here and what the kernels evaluate on
    chip (`ss_blocks`: a size-2 block whose lam carries the SIGN BIT is an oscillator of angular frequency -lam;
    j = 0 is the identity block, lam = -0.0).  `include_dt` / `include_dt2` add the derivative rows of H (:238-245)."""

    def __init__(self, frequency, lengthscale, variance, n_terms=10, include_dt=False, include_dt2=False):
        self.frequency = float(frequency)
        self.lengthscale = float(lengthscale)
        self.variance = float(variance)
        self.n_terms = self.order = int(n_terms)
        self.include_dt, self.include_dt2 = bool(include_dt), bool(include_dt2)
        self._state_space_dim = 2 * (self.n_terms + 1)
        self.input_dim = 1

    def _q2(self):
        from scipy.special import ive          # tfp.math.bessel_ive of the reference (:226)
        J = self.order
        return np.array([1.0] + [2.0] * J) * self.variance * ive(np.arange(J + 1, dtype=np.float64),
                                                                  self.lengthscale ** (-2))

    def to_ss(self, X_spatial=None):
        """(F, L, Qc, H, m_inf, Pinf).  The reference's periodic kernel returns five values (no m_inf, :247), which its
        own `Independent.state_space_representation` (transform.py:400-408, out_dim = 6) cannot stack; the zero
        stationary mean is supplied here so that the kernel composes like every other Markov kernel."""
        J, w = self.order, self.frequency
        j = np.arange(J + 1, dtype=np.float64)
        F = np.kron(np.diag(j), np.array([[0.0, -w], [w, 0.0]]))
        Pinf = np.kron(np.diag(self._q2()), np.eye(2))
        H = np.kron(np.ones([1, J + 1]), np.array([[1.0, 0.0]]))
        if self.include_dt2:
            H = np.vstack([H, np.kron(-j * w, np.array([0.0, 1.0])), np.kron(-j * w, np.array([1.0, 0.0]))])
        elif self.include_dt:
            H = np.vstack([H, np.kron(-j * w, np.array([0.0, 1.0]))])
        return F, np.eye(2 * (J + 1)), np.zeros([2 * (J + 1), 2 * (J + 1)]), H, np.zeros([2 * (J + 1), 1]), Pinf

    def expm(self, dt, X_spatial=None):
        ang = np.arange(self.order + 1, dtype=np.float64) * self.frequency * dt
        return _block_diag([np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) for a in ang])

    def K(self, X1, X2):
        tau = np.abs(X1.reshape(-1, 1) - X2.reshape(1, -1))
        return self.variance * np.exp(-2.0 * np.square(np.sin(self.frequency * tau / 2.0) / self.lengthscale))

    def ss_blocks(self):
        return [(2, -(j * self.frequency)) for j in range(self.order + 1)]     # -(0 * w) = -0.0: identity block


class SumKernel(MarkovKernel):
    """stgp/kernels/kernel.py:134-160: block-diagonal F/Pinf/expm, H = hstack."""

    def __init__(self, k1, k2):
        self.k1, self.k2 = k1, k2

    def state_space_dim(self):
        return self.k1.state_space_dim() + self.k2.state_space_dim()

    state_size = state_space_dim

    def to_ss(self, X_spatial=None):
        a, b = self.k1.to_ss(X_spatial), self.k2.to_ss(X_spatial)
        return (_block_diag([a[0], b[0]]), _block_diag([a[1], b[1]]), _block_diag([a[2], b[2]]),
                np.hstack([a[3], b[3]]), np.vstack([a[4], b[4]]), _block_diag([a[5], b[5]]))

    def expm(self, dt, X_spatial=None):
        return _block_diag([self.k1.expm(dt, X_spatial), self.k2.expm(dt, X_spatial)])

    def K(self, X1, X2):
        return self.k1.K(X1, X2) + self.k2.K(X1, X2)

    def ss_blocks(self):
        a, b = self.k1.ss_blocks(), self.k2.ss_blocks()
        return None if (a is None or b is None) else a + b


def sum_kernels(parts):
    """Left-fold a list of kernels into nested SumKernels (k = k1 + k2 + ...)."""
    out = parts[0]
    for p in parts[1:]:
        out = SumKernel(out, p)
    return out


class SpatioTemporalSeperableKernel(MarkovKernel):
    """stgp/kernels/kernel.py:213-265 with ss_utils.py:41-53: A = I (x) A_t, Pinf = K_s (x) Pinf_t.
    `K_spatial` is the already-evaluated spatial Gram matrix (spatial kernels are outside the path)."""

    def __init__(self, K_temporal, K_spatial):
        self.k1 = K_temporal
        self.Ks = np.asarray(K_spatial, dtype=np.float64)

    def state_space_dim(self):
        return self.k1.state_space_dim()

    def to_ss(self, X_spatial=None):
        F, L, Qc, H, minf, Pinf = self.k1.to_ss()
        Ns = self.Ks.shape[0]
        eye = np.eye(Ns)
        return (np.kron(eye, F), np.kron(eye, L), np.kron(self.Ks, Qc), np.kron(eye, H),
                np.kron(np.ones([Ns, 1]), minf), np.kron(self.Ks, Pinf))

    def expm(self, dt, X_spatial=None):
        return np.kron(np.eye(self.Ks.shape[0]), self.k1.expm(dt))

    def ss_blocks(self):
        return None   # Pinf is dense (K_s (x) Pinf_t): not a block-diagonal stationary stack
