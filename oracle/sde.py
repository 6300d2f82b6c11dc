"""Oracle: LTI-SDE prior API and discretisation (numpy restatement, test infrastructure).

Follows (paths relative to /root/reference/src/lib/stgp/):
  kernels/ss_utils.py:6-53        Matern-3/2 expm + state-space rep, space-time kron rep
  kernels/matern.py:109-177       Matern-5/2 rep + closed-form expm
  kernels/matern.py:269-329       Matern-7/2 rep + closed-form expm
  kernels/kernel.py:134-160       SumKernel block-diagonal stacking
  kernels/kernel.py:207-209       Q_k = Pinf - A_k Pinf A_k^T   (stationary)
  kernels/kernel.py:213-265       SpatioTemporalSeperableKernel (A = I kron A_t, Pinf = K_s kron Pinf_t)
  transforms/transform.py:400-545 Independent: block-diagonal stacking of latents
  transforms/sdes.py:55-172       LTI_SDE / LTI_SDE_Full_State_Obs prior API (m_inf, P_inf, H, expm, Q)
"""
import numpy as np
import scipy.linalg as sla


def block_diag(blocks):
    return sla.block_diag(*blocks)


class Matern32:
    """kernels/ss_utils.py:6-38, kernels/matern.py:53-80 (unit variance) / :13-38 (scaled)."""
    state_dim = 2

    def __init__(self, lengthscale, variance=1.0):
        self.ls = float(lengthscale)
        self.var = float(variance)

    def to_ss(self):
        ls, var = self.ls, self.var
        lam = 3.0 ** 0.5 / ls
        F = np.array([[0.0, 1.0], [-(lam ** 2), -2 * lam]])
        L = np.array([[0.0], [1.0]])
        H = np.array([[1.0, 0.0]])
        Qc = np.array([[12.0 * 3.0 ** 0.5 / ls ** 3.0 * var]])
        minf = np.zeros([2, 1])
        Pinf = np.array([[var, 0.0], [0.0, 3.0 * var / ls ** 2.0]])
        return F, L, Qc, H, minf, Pinf

    def expm(self, dt):
        lam = np.sqrt(3.0) / self.ls
        return np.exp(-dt * lam) * (dt * np.array([[lam, 1.0], [-lam ** 2.0, -lam]]) + np.eye(2))

    def K(self, x1, x2):
        """kernels/matern.py:82-90"""
        r = np.abs(x1[:, None] - x2[None, :]) / self.ls
        s3 = np.sqrt(3.0)
        return self.var * (1.0 + s3 * r) * np.exp(-s3 * r)


class Matern52:
    """kernels/matern.py:109-188 (and the Scaled variant :191-266)."""
    state_dim = 3

    def __init__(self, lengthscale, variance=1.0):
        self.ls = float(lengthscale)
        self.var = float(variance)

    def to_ss(self):
        ls, var = self.ls, self.var
        lam = 5.0 ** 0.5 / ls
        F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0],
                      [-(lam ** 3.0), -3.0 * lam ** 2.0, -3.0 * lam]])
        L = np.array([[0.0], [0.0], [1.0]])
        Qc = np.array([[var * 400.0 * 5.0 ** 0.5 / 3.0 / ls ** 5.0]])
        H = np.array([[1.0, 0.0, 0.0]])
        kappa = 5.0 / 3.0 * var / ls ** 2.0
        minf = np.zeros([3, 1])
        Pinf = np.array([[var, 0.0, -kappa], [0.0, kappa, 0.0],
                         [-kappa, 0.0, 25.0 * var / ls ** 4.0]])
        return F, L, Qc, H, minf, Pinf

    def expm(self, dt):
        lam = np.sqrt(5.0) / self.ls
        dtlam = dt * lam
        M = np.array([
            [lam * (0.5 * dtlam + 1.0), dtlam + 1.0, 0.5 * dt],
            [-0.5 * dtlam * lam ** 2, lam * (1.0 - dtlam), 1.0 - 0.5 * dtlam],
            [lam ** 3 * (0.5 * dtlam - 1.0), lam ** 2 * (dtlam - 3), lam * (0.5 * dtlam - 2.0)],
        ])
        return np.exp(-dtlam) * (dt * M + np.eye(3))

    def K(self, x1, x2):
        r = np.abs(x1[:, None] - x2[None, :]) / self.ls
        s5 = np.sqrt(5.0)
        return self.var * (1.0 + s5 * r + (5.0 / 3.0) * r * r) * np.exp(-s5 * r)


class Matern72:
    """kernels/matern.py:269-341 (ScaledMatern72)."""
    state_dim = 4

    def __init__(self, lengthscale, variance=1.0):
        self.ls = float(lengthscale)
        self.var = float(variance)

    def to_ss(self):
        ls, var = self.ls, self.var
        lam = 7.0 ** 0.5 / ls
        F = np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0],
                      [-lam ** 4.0, -4.0 * lam ** 3.0, -6.0 * lam ** 2.0, -4.0 * lam]])
        L = np.array([[0.0], [0.0], [0.0], [1.0]])
        Qc = np.array([[var * 10976.0 * 7.0 ** 0.5 / 5.0 / ls ** 7.0]])
        H = np.array([[1.0, 0.0, 0.0, 0.0]])
        kappa = 7.0 / 5.0 * var / ls ** 2.0
        kappa2 = 9.8 * var / ls ** 4.0
        minf = np.zeros([4, 1])
        Pinf = np.array([[var, 0.0, -kappa, 0.0], [0.0, kappa, 0.0, -kappa2],
                         [-kappa, 0.0, kappa2, 0.0], [0.0, -kappa2, 0.0, 343.0 * var / ls ** 6.0]])
        return F, L, Qc, H, minf, Pinf

    def expm(self, dt):
        lam = np.sqrt(7.0) / self.ls
        lam2 = lam * lam
        lam3 = lam2 * lam
        dtlam = dt * lam
        dtlam2 = dtlam ** 2
        M = np.array([
            [lam * (1.0 + 0.5 * dtlam + dtlam2 / 6.0), 1.0 + dtlam + 0.5 * dtlam2,
             0.5 * dt * (1.0 + dtlam), dt ** 2 / 6],
            [-dtlam2 * lam ** 2.0 / 6.0, lam * (1.0 + 0.5 * dtlam - 0.5 * dtlam2),
             1.0 + dtlam - 0.5 * dtlam2, dt * (0.5 - dtlam / 6.0)],
            [lam3 * dtlam * (dtlam / 6.0 - 0.5), dtlam * lam2 * (0.5 * dtlam - 2.0),
             lam * (1.0 - 2.5 * dtlam + 0.5 * dtlam2), 1.0 - dtlam + dtlam2 / 6.0],
            [lam2 ** 2 * (dtlam - 1.0 - dtlam2 / 6.0), lam3 * (3.5 * dtlam - 4.0 - 0.5 * dtlam2),
             lam2 * (4.0 * dtlam - 6.0 - 0.5 * dtlam2), lam * (1.5 * dtlam - 3.0 - dtlam2 / 6.0)],
        ])
        return np.exp(-dtlam) * (dt * M + np.eye(4))

    def K(self, x1, x2):
        r = np.abs(x1[:, None] - x2[None, :]) / self.ls
        s7 = np.sqrt(7.0)
        return self.var * (1. + s7 * r + 14. / 5. * r ** 2 + 7. * s7 / 15. * r ** 3) * np.exp(-s7 * r)


class IWP:
    """kernels/wiener.py:60-149 (WienerVelocity / IntegratedWiener): state (f, f', ..., f^(q)), F = shift matrix;
    expm :105-123, Q :125-149 (its own closed form: the process is not stationary), to_ss :90-103 with
    Pinf = stable_state_covariance * I used as the initial covariance."""

    def __init__(self, q=1, variance=1.0, stable_state_covariance=0.0):
        self.q, self.var, self.ssc = int(q), float(variance), float(stable_state_covariance)
        self.state_dim = self.q + 1

    def to_ss(self):
        dim, q = self.q + 1, self.q
        F = np.eye(dim, k=1)
        L = np.hstack([np.zeros(dim - 1), [1.0]])[:, None]
        H = np.hstack([[1.0], np.zeros(q)])[None, :]
        return F, L, np.array([[self.var]]), H, np.zeros([dim, 1]), np.eye(dim) * self.ssc

    def expm(self, dt):
        from math import factorial
        dim = self.q + 1
        return np.array([[dt ** (j - i) / factorial(j - i) if j >= i else 0.0 for j in range(dim)]
                         for i in range(dim)])

    def Q(self, dt, A=None, Pinf=None):
        from math import factorial
        dim, q = self.q + 1, self.q
        return self.var * np.array([[dt ** (2 * q + 1 - i - j) / ((2 * q + 1 - i - j) * factorial(q - i) * factorial(q - j))
                                     for j in range(dim)] for i in range(dim)])


class GenericLTI:
    """A stationary LTI SDE given by (F, H, Pinf) with A = scipy expm(F dt) -- the role played in
    the reference by kernels that call jax.scipy.linalg.expm (kernels/periodic.py:250-253)."""

    def __init__(self, F, H, Pinf, minf=None):
        self.F = np.asarray(F, float)
        self.H = np.asarray(H, float)
        self.Pinf = np.asarray(Pinf, float)
        self.state_dim = self.F.shape[0]
        self.minf = np.zeros([self.state_dim, 1]) if minf is None else np.asarray(minf, float).reshape(-1, 1)

    def to_ss(self):
        d = self.state_dim
        return self.F, np.zeros([d, 1]), np.zeros([1, 1]), self.H, self.minf, self.Pinf

    def expm(self, dt):
        return sla.expm(self.F * dt)


class ApproxPeriodicBN(GenericLTI):
    """kernels/periodic.py:171-253 (`ApproxSDEPeriodic_BN`): the periodic covariance as a sum of n_terms + 1 harmonic
    oscillators (Solin & Sarkka 2014).  F = kron(diag(0..J), [[0, -w], [w, 0]]) (:231), Pinf = kron(diag(q2), I_2) (:234)
    with q2_j = (1 if j == 0 else 2) * variance * ive(j, lengthscale^-2) (:226; tfp's `bessel_ive` is scipy's `ive`),
    H = kron(ones, [1, 0]) (:236), A = expm(F dt) through the generic Pade `expm` (:250-253) as in the reference."""

    def __init__(self, frequency, lengthscale, variance, n_terms=10):
        from scipy.special import ive
        J = int(n_terms)
        q2 = np.array([1.0] + [2.0] * J) * variance * ive(np.arange(J + 1, dtype=float), float(lengthscale) ** (-2))
        F = np.kron(np.diag(np.arange(J + 1, dtype=float)), np.array([[0.0, -frequency], [frequency, 0.0]]))
        Pinf = np.kron(np.diag(q2), np.eye(2))
        H = np.kron(np.ones([1, J + 1]), np.array([[1.0, 0.0]]))
        super().__init__(F, H, Pinf)


class SumKernel:
    """kernels/kernel.py:134-160: block-diagonal F/Pinf/expm, hstacked H."""

    def __init__(self, parts):
        self.parts = list(parts)
        self.state_dim = sum(p.state_dim for p in self.parts)

    def to_ss(self):
        reps = [p.to_ss() for p in self.parts]
        F = block_diag([r[0] for r in reps])
        L = block_diag([r[1] for r in reps])
        Qc = block_diag([r[2] for r in reps])
        H = np.hstack([r[3] for r in reps])
        minf = np.vstack([r[4] for r in reps])
        Pinf = block_diag([r[5] for r in reps])
        return F, L, Qc, H, minf, Pinf

    def expm(self, dt):
        return block_diag([p.expm(dt) for p in self.parts])

    def K(self, x1, x2):
        return sum(p.K(x1, x2) for p in self.parts)


class SpaceTimeSeparable:
    """kernels/kernel.py:213-265 + kernels/ss_utils.py:41-53.

    K_spatial is passed in already evaluated (the spatial kernel itself is outside the hot path,
    SURVEY.md section 2 row 14)."""

    def __init__(self, temporal, K_spatial):
        self.kt = temporal
        self.Ks = np.asarray(K_spatial, float)
        self.Ns = self.Ks.shape[0]
        self.state_dim = self.kt.state_dim * self.Ns

    def to_ss(self):
        F, L, Qc, H, minf, Pinf = self.kt.to_ss()
        eye = np.eye(self.Ns)
        return (np.kron(eye, F), np.kron(eye, L), np.kron(self.Ks, Qc), np.kron(eye, H),
                np.kron(np.ones([self.Ns, 1]), minf), np.kron(self.Ks, Pinf))

    def expm(self, dt):
        return np.kron(np.eye(self.Ns), self.kt.expm(dt))


class LTI_SDE:
    """transforms/sdes.py:18-97 over transforms/transform.py:400-545 (`Independent` of Q latents).

    `latents` is a list of kernels; the joint state stacks them block-diagonally
    (time-latent-space-state ordering, computation/filters/kalman_filter.py:4-14)."""

    def __init__(self, latents, m_init=None):
        self.latents = list(latents)
        self.m_init = None if m_init is None else np.asarray(m_init, float).reshape(-1, 1)
        self.state_dim = sum(k.state_dim for k in self.latents)

    def state_space_representation(self):
        reps = [k.to_ss() for k in self.latents]
        F = block_diag([r[0] for r in reps])
        L = block_diag([r[1] for r in reps])
        Qc = block_diag([r[2] for r in reps])
        H = block_diag([r[3] for r in reps])
        minf = np.vstack([r[4] for r in reps])
        Pinf = block_diag([r[5] for r in reps])
        return F, L, Qc, H, minf, Pinf

    def m_inf(self):
        if self.m_init is not None:
            return self.m_init
        return self.state_space_representation()[4]

    def P_inf(self):
        return self.state_space_representation()[5]

    def H(self):
        return self.state_space_representation()[3]

    def expm(self, dt):
        return block_diag([k.expm(dt) for k in self.latents])

    def Q(self, dt, A, Pinf):
        # kernel.py:207-209 applied per latent block then re-stacked (transform.py:499-545);
        # with block-diagonal A and Pinf that equals the dense expression below.  Kernels with their own Q
        # (the integrated Wiener process, wiener.py:125-149) supply their block.
        if any(hasattr(k, "Q") for k in self.latents):
            out, off = [], 0
            for k in self.latents:
                n = k.state_dim
                sl = slice(off, off + n)
                out.append(k.Q(dt, A[sl, sl], Pinf[sl, sl]) if hasattr(k, "Q")
                           else Pinf[sl, sl] - A[sl, sl] @ Pinf[sl, sl] @ A[sl, sl].T)
                off += n
            return block_diag(out)
        return Pinf - A @ Pinf @ A.T


class LTI_SDE_Full_State_Obs(LTI_SDE):
    """transforms/sdes.py:99-172 with overwrite_H=True, Ns=1, ds=1: H observes every state
    component of every latent (identity up to the latent-df ordering, which for Ns=ds=1 is the
    identity permutation), optionally only `keep_dims` of each latent (sdes.py:174-190)."""

    def __init__(self, latents, keep_dims=None):
        super().__init__(latents)
        self.keep_dims = keep_dims

    def H(self):
        rows = []
        off = 0
        for k in self.latents:
            d = k.state_dim
            keep = range(d) if self.keep_dims is None else self.keep_dims
            for j in keep:
                e = np.zeros(self.state_dim)
                e[off + j] = 1.0
                rows.append(e)
            off += d
        return np.array(rows)


def lyapunov_residual(F, L, Qc, Pinf):
    """F Pinf + Pinf F^T + L Qc L^T  (zero for a correct stationary covariance; SURVEY 8c pin iii)."""
    return F @ Pinf + Pinf @ F.T + L @ Qc @ L.T
