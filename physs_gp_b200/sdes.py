"""Host-side mirror of the prior API the filter consumes:
  stgp/transforms/transform.py:400-545  Independent (block-diagonal stack of Q latent GPs)
  stgp/transforms/sdes.py:18-97         LTI_SDE        (m_inf, P_inf, H, expm, Q, state_space_representation)
  stgp/transforms/sdes.py:99-190        LTI_SDE_Full_State_Obs(_With_Mask)  (H observes the derivative state)
Same method names and argument order, numpy on the host (these are d x d, T-independent objects).

`BatchedMaternSDE` is new: B independent series with per-series hyper-parameters, evaluated
vectorised -- the reference has no batch axis.
"""
import numpy as np

from .kernels import _block_diag, Matern12, Matern32, Matern52, Matern72


class GP:
    """Minimal stand-in for a latent `GP(kernel)` node (stgp/models/gp.py:5-10)."""

    def __init__(self, kernel):
        self.kernel = kernel


class Independent:
    def __init__(self, latents):
        self.parent = [l if hasattr(l, "kernel") else GP(l) for l in latents]

    @property
    def output_dim(self):
        return len(self.parent)

    def state_space_dim(self):
        return [l.kernel.state_space_dim() for l in self.parent]

    def state_space_representation(self, X_s=None):
        reps = [l.kernel.to_ss(X_s) for l in self.parent]
        F, L, Qc, H, minf, Pinf = [[r[i] for r in reps] for i in range(6)]
        return (_block_diag(F), _block_diag(L), _block_diag(Qc), _block_diag(H),
                np.vstack(minf), _block_diag(Pinf))

    def expm(self, dt, X_s=None):
        return _block_diag([l.kernel.expm(dt, X_s) for l in self.parent])

    def Q(self, dt_k, A_k, P_inf, X_spatial=None):
        # transform.py:499-545: per-latent kernel.Q on the diagonal blocks, re-stacked
        out, off = [], 0
        for l in self.parent:
            n = l.kernel.to_ss(X_spatial)[0].shape[0]
            sl = slice(off, off + n)
            out.append(l.kernel.Q(dt_k, A_k[sl, sl], P_inf[sl, sl], X_spatial=X_spatial))
            off += n
        return _block_diag(out)

    def ss_blocks(self):
        blocks = []
        for l in self.parent:
            b = l.kernel.ss_blocks() if hasattr(l.kernel, "ss_blocks") else None
            if b is None:
                return None
            blocks += b
        return blocks

    def iwp_blocks(self):
        blocks = []
        for l in self.parent:
            b = l.kernel.iwp_blocks() if hasattr(l.kernel, "iwp_blocks") else None
            if b is None:
                return None
            blocks += b
        return blocks


class LTI_SDE:
    """stgp/transforms/sdes.py:18-97."""

    def __init__(self, gp, m_init=None):
        self.gp = gp if isinstance(gp, Independent) else Independent([gp])
        self.m_init = None if m_init is None else np.reshape(np.asarray(m_init, np.float64), [-1, 1])

    def state_space_dim(self):
        return self.gp.state_space_dim()

    def state_space_representation(self, X_s=None, dt=None, t=None):
        return self.gp.state_space_representation(X_s)

    def H(self, x=None, X_s=None, t=None):
        return self.gp.state_space_representation(X_s)[3]

    def P_inf(self, x=None, X_s=None, t=None):
        return self.gp.state_space_representation(X_s)[5]

    def m_inf(self, x=None, X_s=None, t=None):
        if self.m_init is not None:
            return self.m_init
        return self.gp.state_space_representation(X_s)[4]

    def expm(self, X_s, t):
        return self.gp.expm(t, X_s)

    def Q(self, dt_k, A_k, P_inf, X_spatial=None):
        return self.gp.Q(dt_k, A_k, P_inf, X_spatial=X_spatial)

    def ss_blocks(self):
        return self.gp.ss_blocks()

    def iwp_blocks(self):
        return self.gp.iwp_blocks()


class LTI_SDE_Full_State_Obs(LTI_SDE):
    """stgp/transforms/sdes.py:99-172 for temporal models (Ns = ds = 1, overwrite_H=True): H selects
    `keep_dims` of every latent's state (all of it by default), i.e. a row-subset of the identity."""

    def __init__(self, gp, keep_dims=None):
        super().__init__(gp)
        self.keep_dims = None if keep_dims is None else list(keep_dims)

    def H(self, x=None, X_s=None, t=None):
        dims = self.gp.state_space_dim()
        d = sum(dims)
        rows, off = [], 0
        for n in dims:
            for j in (range(n) if self.keep_dims is None else self.keep_dims):
                e = np.zeros(d)
                e[off + j] = 1.0
                rows.append(e)
            off += n
        return np.array(rows)


LTI_SDE_Full_State_Obs_With_Mask = LTI_SDE_Full_State_Obs   # sdes.py:174-190 (keep_dims mandatory)

_KINDS = {1: Matern12, 2: Matern32, 3: Matern52, 4: Matern72}


class PointResidual:
    """One point-wise collocation residual on the derivative-augmented state x:
        g(x, t_k) = w . x + sum_q coef_q * phi_q(x[idx_q]) + forcing[k],   phi in {'sin', 'cos', 'square', 'cube'}
    plus bilinear terms ('prod', (i, j), coef) = coef * x[i] * x[j]
    -- the form of the residuals the reference ships as `PDE.forward_g` (transforms/pdes.py: Pendulum1D :482-528,
    DampedPendulum1D :530-597, SimpleODE :424-480, the u^3 - u reaction of Allen-Cahn :700-811, and with the
    bilinear terms the components of LotkaVolterra :912-1090 and LorenzSystem :818-910).  Its Jacobian,
    which the reference takes with jax.jacfwd (`PDE.jac`, :236-242), is evaluated on chip."""

    def __init__(self, w, terms=(), forcing=None):
        self.w = np.asarray(w, np.float64)
        self.terms = [(str(k), (int(i[0]) | (int(i[1]) << 8)) if isinstance(i, (tuple, list)) else int(i), float(c))
                      for k, i, c in terms]
        self.forcing = None if forcing is None else np.asarray(forcing, np.float64)


class PDE:
    """Collocation prior: PDE[LTI_SDE[GP]] of the reference (transforms/pdes.py:227-245) -- an LTI_SDE `parent`
    constrained at every time step by pseudo-observations of point-wise residuals (EKF-style collocation,
    kalman_filter.py:340-427).  `filter_type='b200'` runs it with `physs_kf_filter_colloc_f64`; the smoother is the
    parent's (rts_smoother.py:108-150).

    residuals: 1 or 2 PointResidual (up to 3 over three latents: systems of ODEs such as `LorenzSystem`, two latents for
    `LotkaVolterra` -- independent Matern-3/2 latents, one observed output each); psuedo_observations: one value per residual, 0 (collocate) or NaN (off);
    boundary_conditions: [Nt, m] array with NaN where there is no boundary observation, or None; observe_data as in
    the reference (PDE.observe_data, default False there; True here only when asked)."""

    def __init__(self, parent, residuals, psuedo_observations=None, boundary_conditions=None, observe_data=False):
        self.parent = parent
        self.residuals = list(residuals)
        if not 1 <= len(self.residuals) <= 3:
            raise ValueError("1 to 3 collocation residuals per time step")
        self._pseudo = (np.zeros(len(self.residuals)) if psuedo_observations is None
                        else np.asarray(psuedo_observations, np.float64).reshape(len(self.residuals)))
        self.boundary_conditions = boundary_conditions
        self.observe_data = observe_data

    def psuedo_observations(self, X_s=None):
        return self._pseudo[:, None]

    # the LTI part is the parent's (pdes.py:233-234 and the filter's `sde_prior = model.parent`)
    def m_inf(self, x=None, X_s=None, t=None):
        return self.parent.m_inf(x, X_s, t)

    def P_inf(self, x=None, X_s=None, t=None):
        return self.parent.P_inf(x, X_s, t)

    def H(self, x=None, X_s=None, t=None):
        return self.parent.H(x, X_s, t)

    def expm(self, X_s, t):
        return self.parent.expm(X_s, t)

    def Q(self, dt_k, A_k, P_inf, X_spatial=None):
        return self.parent.Q(dt_k, A_k, P_inf, X_spatial)

    def ss_blocks(self):
        return self.parent.ss_blocks() if hasattr(self.parent, "ss_blocks") else None

    def iwp_blocks(self):
        return self.parent.iwp_blocks() if hasattr(self.parent, "iwp_blocks") else None


class BatchedMaternSDE:
    """B independent series, each a stack of `nblk` Matern-(s-1/2) blocks of equal size s with its own
    lengthscales/variances [B, nblk]; `sum_blocks=True` observes the SUM of the blocks (SumKernel,
    H = hstack) and False observes each block separately (Independent, H = block-diag).
    `full_state_obs=True` gives H = I (LTI_SDE_Full_State_Obs)."""

    def __init__(self, block_size, lengthscales, variances=None, sum_blocks=True, full_state_obs=False):
        self.s = int(block_size)
        self.ls = np.atleast_2d(np.asarray(lengthscales, np.float64))
        self.var = np.ones_like(self.ls) if variances is None else np.broadcast_to(
            np.atleast_2d(np.asarray(variances, np.float64)), self.ls.shape).copy()
        self.B, self.nblk = self.ls.shape
        self.d = self.s * self.nblk
        self.sum_blocks = sum_blocks
        self.full_state_obs = full_state_obs

    def lam(self):
        return np.sqrt(2.0 * self.s - 1.0) / self.ls          # [B, nblk]

    def P_inf(self):
        """[B, d, d] block-diagonal stationary covariances (closed forms of kernels.py, vectorised)."""
        lam, v, s = self.lam(), self.var, self.s
        blk = np.zeros([self.B, self.nblk, s, s])
        if s == 1:
            blk[..., 0, 0] = v
        elif s == 2:
            blk[..., 0, 0] = v
            blk[..., 1, 1] = lam ** 2 * v
        elif s == 3:
            k = lam ** 2 * v / 3.0
            blk[..., 0, 0] = v
            blk[..., 1, 1] = k
            blk[..., 0, 2] = blk[..., 2, 0] = -k
            blk[..., 2, 2] = lam ** 4 * v
        elif s == 4:
            k1, k2 = lam ** 2 * v / 5.0, lam ** 4 * v / 5.0
            blk[..., 0, 0] = v
            blk[..., 1, 1] = k1
            blk[..., 2, 2] = k2
            blk[..., 3, 3] = lam ** 6 * v
            blk[..., 0, 2] = blk[..., 2, 0] = -k1
            blk[..., 1, 3] = blk[..., 3, 1] = -k2
        else:
            raise ValueError("block size must be 1..4")
        P = np.zeros([self.B, self.d, self.d])
        for b in range(self.nblk):
            P[:, b * s:(b + 1) * s, b * s:(b + 1) * s] = blk[:, b]
        return P

    def m_inf(self):
        return np.zeros([1, self.d])

    def hyper_grads(self, glam, gPinf):
        """Chain gradients with respect to (lam [B, nblk], Pinf [B, d, d]) -- what the filter's reverse pass
        returns, with the P0 = Pinf contribution already added to gPinf -- to the kernels' own parameters:
        lam = sqrt(2 s - 1) / lengthscale and, in every closed form above, Pinf[i][j] = c_ij * variance *
        lam^(i + j) inside a block.  Returns (d/d lengthscale, d/d variance), each [B, nblk]."""
        glam, gPinf = np.asarray(glam, np.float64), np.asarray(gPinf, np.float64)
        lam, P, s = self.lam(), self.P_inf(), self.s
        powers = np.add.outer(np.arange(s), np.arange(s)).astype(np.float64)        # i + j
        g_ls, g_var = np.zeros_like(lam), np.zeros_like(lam)
        for b in range(self.nblk):
            sl = slice(b * s, (b + 1) * s)
            gp, pb = gPinf[:, sl, sl], P[:, sl, sl]
            dlam = glam[:, b] + np.sum(gp * pb * powers, axis=(1, 2)) / lam[:, b]
            g_ls[:, b] = -dlam * lam[:, b] / self.ls[:, b]
            g_var[:, b] = np.sum(gp * pb, axis=(1, 2)) / self.var[:, b]
        return g_ls, g_var

    def H(self):
        if self.full_state_obs:
            return np.eye(self.d)
        if self.sum_blocks:
            h = np.zeros([1, self.d])
            h[0, ::self.s] = 1.0
            return h
        H = np.zeros([self.nblk, self.d])
        for b in range(self.nblk):
            H[b, b * self.s] = 1.0
        return H

    def series(self, b):
        """The b-th series as an ordinary (reference-shaped) prior object."""
        from .kernels import sum_kernels
        ks = [_KINDS[self.s](self.ls[b, i], self.var[b, i]) for i in range(self.nblk)]
        if self.full_state_obs:
            return LTI_SDE_Full_State_Obs(Independent(ks))
        if self.sum_blocks:
            return LTI_SDE(Independent([sum_kernels(ks)]))
        return LTI_SDE(Independent(ks))
