// physs_pde.cu -- PDE-collocation expected log-likelihood and the Gauss-Newton site curvature for the
// damped-oscillator residual (BASELINE config 3), one thread per site block, flat over N = B*T blocks.
//
// Reference: the PHYSS-GP model stacks [OutputMap (observe x), DampedPendulum1D residual] as a MultiOutput
// prior transform (zoo/sde_diff.py:757-763, transforms/pdes.py:530-597: r(f) = f_tt + (g/l) sin f + b f_t)
// with Gaussian likelihoods on both outputs (the residual is "observed" as 0 with collocation noise).  Its
// CVI step takes dELL/dm by jax.grad of a Monte-Carlo ELL (cvi_nat_grad.py:381-383, approximators.py:16-58)
// and replaces dELL/dS by the Laplace-Gauss-Newton delta-u curvature
//   dS := 1/2 sum_p mask_p J_p^T (-1/sigma_p^2) J_p,   J_p = d T_p(u)/du at u = q_mu
// (cvi_hessian_approximations.py:333-431,483-486,542-577; enforce_psd_type='laplace_gauss_newton_delta_u',
// E/pendulum/models/m_stgp.py:243).
//
// Here the expectation is evaluated in CLOSED FORM instead of Monte-Carlo: with u ~ N(m, S), l = u2 + b u1 - y_c
// (Gaussian) and s = sin u0,
//   E[(l + a s)^2] = E[l^2] + 2 a (E[l] E[s] + cov(l, u0) E[cos u0]) + a^2 E[s^2]          (Stein's lemma)
//   E[sin u0] = sin(m0) e^{-S00/2},  E[cos u0] = cos(m0) e^{-S00/2},  E[sin^2 u0] = (1 - cos(2 m0) e^{-2 S00}) / 2,
// which is exact, deterministic and differentiable analytically (the MC estimate is parity-unpinned anyway:
// it depends on an objax PRNG stream).
#include "physs_core.cuh"
#include "physs_internal.h"

namespace physs {

struct PendArgs {
  int64_t N; int D; int i0, i1, i2;
  const double* qm; const double* qS; const double* y;   // y [N, 2]: (observation of x, collocation target); NaN = absent
  double a, b, var_obs, var_col;
  int gauss_newton;
  double* ell; double* dm; double* dS;
};

__global__ void __launch_bounds__(128) pendulum_ell_kernel(const PendArgs p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int D = p.D, i0 = p.i0, i1 = p.i1, i2 = p.i2;
  const double* __restrict__ qm = p.qm + n * D;
  const double* __restrict__ qS = p.qS + n * (int64_t)D * D;
  const double m0 = qm[i0], m1 = qm[i1], m2 = qm[i2];
  const double S00 = qS[i0 * D + i0], S11 = qS[i1 * D + i1], S22 = qS[i2 * D + i2];
  const double S01 = 0.5 * (qS[i0 * D + i1] + qS[i1 * D + i0]);
  const double S02 = 0.5 * (qS[i0 * D + i2] + qS[i2 * D + i0]);
  const double S12 = 0.5 * (qS[i1 * D + i2] + qS[i2 * D + i1]);
  const double yo = p.y[n * 2], yc = p.y[n * 2 + 1];
  const bool has_o = !(yo != yo), has_c = !(yc != yc);
  const double a = p.a, b = p.b;
  double ell = 0.0;
  double g0 = 0.0, g1 = 0.0, g2 = 0.0;                       // dELL/dm on (i0, i1, i2)
  double h00 = 0.0, h11 = 0.0, h22 = 0.0, h01 = 0.0, h02 = 0.0, h12 = 0.0;   // dELL/dS entries
  if (has_o) {
    const double r = 1.0 / p.var_obs, e = yo - m0;
    ell += -0.5 * (kLog2Pi + log(p.var_obs) + (e * e + S00) * r);
    g0 += e * r;
    h00 += -0.5 * r;
  }
  if (has_c) {
    const double r = 1.0 / p.var_col;
    const double E = exp(-0.5 * S00), E4 = exp(-2.0 * S00);
    double sm, cm;
    sincos(m0, &sm, &cm);
    const double mul = m2 + b * m1 - yc;                      // E[l]
    const double vl = S22 + 2.0 * b * S12 + b * b * S11;      // var l
    const double c = S02 + b * S01;                           // cov(l, u0)
    const double c2m = cm * cm - sm * sm, s2m = 2.0 * sm * cm;
    const double A1 = mul * mul + vl;
    const double A2 = (mul * sm + c * cm) * E;
    const double A3 = 0.5 * (1.0 - c2m * E4);
    const double Rr = A1 + 2.0 * a * A2 + a * a * A3;
    ell += -0.5 * (kLog2Pi + log(p.var_col) + Rr * r);
    const double k = -0.5 * r;                                // dELL = k * dRr
    const double t = 2.0 * mul + 2.0 * a * sm * E;
    g0 += k * (2.0 * a * (mul * cm - c * sm) * E + a * a * s2m * E4);
    g1 += k * b * t;
    g2 += k * t;
    if (p.gauss_newton) {
      // J = [a cos(m0), b, 1] at u = q_mu (delta-u, delta-f): dS = -1/2 J^T J / var_col
      const double j0 = a * cm;
      h00 += k * j0 * j0; h11 += k * b * b; h22 += k;
      h01 += k * j0 * b;  h02 += k * j0;    h12 += k * b;
    } else {
      h00 += k * (-a * (mul * sm + c * cm) * E + a * a * c2m * E4);
      h11 += k * b * b; h22 += k;
      h01 += k * a * b * cm * E; h02 += k * a * cm * E; h12 += k * b;
    }
  }
  if (p.ell) p.ell[n] = ell;
  if (p.dm) {
    double* __restrict__ dm = p.dm + n * D;
    for (int i = 0; i < D; ++i) dm[i] = 0.0;
    dm[i0] = g0; dm[i1] = g1; dm[i2] = g2;
  }
  if (p.dS) {
    double* __restrict__ dS = p.dS + n * (int64_t)D * D;
    for (int i = 0; i < D * D; ++i) dS[i] = 0.0;
    dS[i0 * D + i0] = h00; dS[i1 * D + i1] = h11; dS[i2 * D + i2] = h22;
    dS[i0 * D + i1] = h01; dS[i1 * D + i0] = h01;
    dS[i0 * D + i2] = h02; dS[i2 * D + i0] = h02;
    dS[i1 * D + i2] = h12; dS[i2 * D + i1] = h12;
  }
}

}  // namespace physs

extern "C" {

int physs_cvi_ell_pendulum_f64(void* stream, int64_t N, int32_t D, int32_t i0, int32_t i1, int32_t i2,
                               const double* q_mu, const double* q_var, const double* y,
                               double g_over_l, double damping, double var_obs, double var_col,
                               int32_t gauss_newton, double* ell_out, double* dm_out, double* dS_out) {
  using namespace physs;
  if (N < 0 || D < 3) return set_error(PHYSS_ERR_BAD_ARG, "pendulum ell: need N >= 0 and D >= 3");
  if (N == 0) return PHYSS_OK;
  if (!q_mu || !q_var || !y) return set_error(PHYSS_ERR_BAD_ARG, "pendulum ell: null required pointer");
  const int idx[3] = {i0, i1, i2};
  for (int k = 0; k < 3; ++k)
    if (idx[k] < 0 || idx[k] >= D) return set_error(PHYSS_ERR_BAD_ARG, "pendulum ell: state index out of range");
  if (i0 == i1 || i0 == i2 || i1 == i2) return set_error(PHYSS_ERR_BAD_ARG, "pendulum ell: state indices must differ");
  if (!(var_obs > 0.0) || !(var_col > 0.0)) return set_error(PHYSS_ERR_BAD_ARG, "pendulum ell: variances must be positive");
  PendArgs a{N, D, i0, i1, i2, q_mu, q_var, y, g_over_l, damping, var_obs, var_col, gauss_newton,
             ell_out, dm_out, dS_out};
  pendulum_ell_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "pendulum_ell_kernel launch");
}

}

// ------------------------------------------------------------------------------------------------
// Generic Gauss-Newton site curvature from caller-supplied Jacobians (any prior transform): the reference
// obtains J_p = d T_p(u) / du per time step by jax.jacfwd (cvi_hessian_approximations.py:358-368) and forms
//   G = sum_p mask_p J_p^T (-Lambda_p^-1) J_p,     dS = 1/2 G      (:380-431, 483-486, 574)
// with Lambda_p the conditional likelihood variance (Laplace-Gauss-Newton) and the mask from missing data.
// One thread per site block.
namespace physs {

struct GnArgs {
  int64_t N; int D, P;
  const double* J;        // [N, P, D]
  const double* var;      // [., P] likelihood variances, var_stride elements between blocks (0 = shared)
  int64_t var_stride;
  const double* y;        // [N, P] data, NaN = missing (NULL = all observed)
  double* dS;             // [N, D, D]
};

__global__ void __launch_bounds__(128) gauss_newton_kernel(const GnArgs p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  const int D = p.D, P = p.P;
  const double* __restrict__ J = p.J + n * (int64_t)P * D;
  const double* __restrict__ var = p.var + n * p.var_stride;
  double* __restrict__ dS = p.dS + n * (int64_t)D * D;
  for (int i = 0; i < D; ++i) {
    for (int j = i; j < D; ++j) {
      double acc = 0.0;
      for (int a = 0; a < P; ++a) {
        const bool obs = p.y ? !(p.y[n * P + a] != p.y[n * P + a]) : true;
        if (obs) acc = fma(J[a * D + i] / var[a], J[a * D + j], acc);
      }
      dS[i * D + j] = -0.5 * acc;
      dS[j * D + i] = -0.5 * acc;
    }
  }
}

}  // namespace physs

extern "C" {

int physs_cvi_gauss_newton_f64(void* stream, int64_t N, int32_t D, int32_t P, const double* J, const double* var,
                               int64_t var_stride, const double* y, double* dS_out) {
  using namespace physs;
  if (N < 0 || D < 1 || P < 1) return set_error(PHYSS_ERR_BAD_ARG, "gauss-newton: bad sizes");
  if (N == 0) return PHYSS_OK;
  if (!J || !var || !dS_out) return set_error(PHYSS_ERR_BAD_ARG, "gauss-newton: null required pointer");
  GnArgs a{N, D, P, J, var, var_stride, y, dS_out};
  gauss_newton_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "gauss_newton_kernel launch");
}

}
