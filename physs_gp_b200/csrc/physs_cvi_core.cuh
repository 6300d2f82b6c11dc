// physs_cvi_core.cuh -- per-site CVI algebra in registers (D <= 4), PHYSS_HD so that tests can check
// it on the host (see physs_core.cuh for the convention).
//
// Reference semantics (paths relative to /root/reference/src/lib/stgp/):
//   computation/natural_gradients/exponential_family_transforms.py:25-42,70-83  theta <-> lambda (ng_jitter)
//   computation/natural_gradients/cvi_nat_grad.py:47-87                         cvi_block_update
//   computation/elbos/expected_log_likelihoods.py:90-117                        block-Gaussian ELL
//   computation/general.py:9-26                                                 Poisson / Bernoulli log-liks
//   computation/integrals/samples.py:67-90                                      Gauss-Hermite (intent)
// The reference differentiates the ELL with jax.grad (cvi_nat_grad.py:381-383); the closed forms are
//   Gaussian:  dm = W^T R^-1 (y - W m),  dS = -1/2 W^T R^-1 W   (missing rows/cols dropped)
//   scalar non-Gaussian site f = w.u:  dm = w E[l'(f)],  dS = w w^T 1/2 E[l''(f)]
#pragma once
#include "physs_core.cuh"

namespace physs {

enum CviLik { CVI_LIK_GAUSS = 0, CVI_LIK_POISSON_EXP = 1, CVI_LIK_BERNOULLI_PROBIT = 2, CVI_LIK_GIVEN = 3 };

constexpr double kInvSqrt2Pi = 0.39894228040143267793994605993438;
constexpr double kInvSqrt2 = 0.70710678118654752440084436210485;

// l(f), l'(f), l''(f) of the scalar likelihoods (general.py:9-26).
// `c0` = y log(binsize) - lgamma(y + 1): the part of l(f) that does not depend on f, evaluated ONCE per site by
// the caller (lgamma inside the quadrature loop cost more than the exp it sits next to).
PHYSS_HD double poisson_exp_const(double y, double binsize) { return y * log(binsize) - lgamma(y + 1.0); }
// The same constant with log(binsize) hoisted out of the kernel and lgamma(y + 1) read from a table of log-factorials
// for the integer counts 0..255 (lgamma is ~250 fp64 instructions with branches: more than the rest of the closed-form
// Poisson site together).  logfact == nullptr, or a count outside the table: the plain evaluation.
constexpr int kLogFactN = 256;
PHYSS_HD double poisson_exp_const_tab(double y, double log_binsize, const double* logfact) {
  const bool tab = logfact != nullptr && y >= 0.0 && y < (double)kLogFactN && y == floor(y);
  return y * log_binsize - (tab ? logfact[(int)y] : lgamma(y + 1.0));
}
PHYSS_HD void poisson_exp_terms(double y, double f, double binsize, double c0, double& l, double& d1, double& d2) {
  const double lam = exp(f) * binsize;
  l = fma(y, f, c0) - lam;
  d1 = y - lam;
  d2 = -lam;
}

PHYSS_HD void bernoulli_probit_terms(double y, double f, double& l, double& d1, double& d2) {
  const double p = 0.5 * erfc(-f * kInvSqrt2);
  const double pdf = exp(-0.5 * f * f) * kInvSqrt2Pi;
  const double a = p + 1e-5, b = 1.0 - p + 1e-5;   // the reference's +1e-5 inside both logs
  l = y * log(a) + (1.0 - y) * log(b);
  const double ra = 1.0 / a, rb = 1.0 / b;
  d1 = y * pdf * ra - (1.0 - y) * pdf * rb;
  const double dpdf = -f * pdf;
  d2 = y * (dpdf * ra - pdf * pdf * ra * ra) - (1.0 - y) * (dpdf * rb + pdf * pdf * rb * rb);
}

// Inverse of an SPD matrix through its (jittered) Cholesky factor: Ainv = (A + jit I)^-1.
template <int N>
PHYSS_HD void spd_inverse(const double (&A)[N][N], double jit, double (&Ainv)[N][N]) {
  double Aj[N][N], L[N][N], rd[N];
  PHYSS_UNROLL
  for (int i = 0; i < N; ++i) {
    PHYSS_UNROLL
    for (int j = 0; j < N; ++j) Aj[i][j] = A[i][j] + (i == j ? jit : 0.0);
  }
  chol_lower<N>(Aj, L, rd);
  PHYSS_UNROLL
  for (int c = 0; c < N; ++c) {
    double x[N];
    PHYSS_UNROLL
    for (int i = 0; i < N; ++i) x[i] = (i == c) ? 1.0 : 0.0;
    chol_solve_vec<N>(L, rd, x);
    PHYSS_UNROLL
    for (int i = 0; i < N; ++i) Ainv[i][c] = x[i];
  }
}

// Block-Gaussian expected log-likelihood  log N(y | f_mu, R) - 1/2 tr(R^-1 f_var)  with NaN masking
// (expected_log_likelihoods.py:90-117), and optionally R^-1 restricted to the observed entries.
template <int P>
PHYSS_HD double gauss_ell(const double (&y)[P], const double (&R)[P][P], const double (&fmu)[P],
                          const double (&fvar)[P][P], double (&Rinv)[P][P], double (&err)[P]) {
  bool obs[P];
  int nobs = 0;
  PHYSS_UNROLL
  for (int a = 0; a < P; ++a) { obs[a] = !(y[a] != y[a]); nobs += obs[a] ? 1 : 0; }
  double Rm[P][P], L[P][P], rd[P];
  PHYSS_UNROLL
  for (int a = 0; a < P; ++a) {
    PHYSS_UNROLL
    for (int b = 0; b < P; ++b) Rm[a][b] = (obs[a] && obs[b]) ? R[a][b] : (a == b ? 1.0 : 0.0);
    err[a] = obs[a] ? (y[a] - fmu[a]) : 0.0;
  }
  chol_lower<P>(Rm, L, rd);
  double logdet = 0.0;
  PHYSS_UNROLL
  for (int a = 0; a < P; ++a) logdet += log(L[a][a] * L[a][a]);
  // R^-1 (masked-to-identity), then zero the missing rows/cols for the gradient
  double tr = 0.0, mahal = 0.0;
  PHYSS_UNROLL
  for (int c = 0; c < P; ++c) {
    double x[P];
    PHYSS_UNROLL
    for (int i = 0; i < P; ++i) x[i] = (i == c) ? 1.0 : 0.0;
    chol_solve_vec<P>(L, rd, x);
    PHYSS_UNROLL
    for (int i = 0; i < P; ++i) {
      const bool keep = obs[i] && obs[c];
      Rinv[i][c] = keep ? x[i] : 0.0;
      tr = fma(keep ? x[i] : 0.0, fvar[c][i], tr);
    }
  }
  PHYSS_UNROLL
  for (int a = 0; a < P; ++a) {
    double t = 0.0;
    PHYSS_UNROLL
    for (int b = 0; b < P; ++b) t = fma(Rinv[a][b], err[b], t);
    mahal = fma(err[a], t, mahal);
  }
  return -0.5 * ((double)nobs * kLog2Pi + logdet + mahal + tr);
}

// ELL and its gradients w.r.t. the site-block marginal (qm, qS) for P outputs f = W u.
//   LIK == GAUSS: noise [P][P];  POISSON/BERNOULLI: independent scalar sites, K-point Gauss-Hermite
//   (nodes ghx, weights ghw already divided by sqrt(pi)); lik_param = Poisson binsize.
template <int D, int P, int LIK>
PHYSS_HD double cvi_ell_grads(const double (&qm)[D], const double (&qS)[D][D], const double (&y)[P],
                              const double (&W)[P][D], const double (&noise)[P][P], double lik_param,
                              int K, const double* ghx, const double* ghw, double (&dm)[D],
                              double (&dS)[D][D], bool want_ell = true, double log_param = 0.0,
                              const double* logfact = nullptr, bool have_log = false) {
  double fmu[P], WS[P][D];
  PHYSS_UNROLL
  for (int a = 0; a < P; ++a) {
    double t = 0.0;
    PHYSS_UNROLL
    for (int k = 0; k < D; ++k) t = fma(W[a][k], qm[k], t);
    fmu[a] = t;
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) {
      double u = 0.0;
      PHYSS_UNROLL
      for (int k = 0; k < D; ++k) u = fma(W[a][k], qS[k][j], u);
      WS[a][j] = u;
    }
  }
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    dm[i] = 0.0;
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) dS[i][j] = 0.0;
  }
  double ell = 0.0;
  if (LIK == CVI_LIK_GAUSS) {
    double fvar[P][P], Rinv[P][P], err[P];
    PHYSS_UNROLL
    for (int a = 0; a < P; ++a) {
      PHYSS_UNROLL
      for (int b = 0; b < P; ++b) {
        double t = 0.0;
        PHYSS_UNROLL
        for (int k = 0; k < D; ++k) t = fma(WS[a][k], W[b][k], t);
        fvar[a][b] = t;
      }
    }
    ell = gauss_ell<P>(y, noise, fmu, fvar, Rinv, err);
    double Re[P], RW[P][D];
    PHYSS_UNROLL
    for (int a = 0; a < P; ++a) {
      double t = 0.0;
      PHYSS_UNROLL
      for (int b = 0; b < P; ++b) t = fma(Rinv[a][b], err[b], t);
      Re[a] = t;
      PHYSS_UNROLL
      for (int j = 0; j < D; ++j) {
        double u = 0.0;
        PHYSS_UNROLL
        for (int b = 0; b < P; ++b) u = fma(Rinv[a][b], W[b][j], u);
        RW[a][j] = u;
      }
    }
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      double t = 0.0;
      PHYSS_UNROLL
      for (int a = 0; a < P; ++a) t = fma(W[a][i], Re[a], t);
      dm[i] = t;
      PHYSS_UNROLL
      for (int j = 0; j < D; ++j) {
        double u = 0.0;
        PHYSS_UNROLL
        for (int a = 0; a < P; ++a) u = fma(W[a][i], RW[a][j], u);
        dS[i][j] = -0.5 * u;
      }
    }
  } else {
    PHYSS_UNROLL
    for (int a = 0; a < P; ++a) {
      double fv = 0.0;
      PHYSS_UNROLL
      for (int k = 0; k < D; ++k) fv = fma(WS[a][k], W[a][k], fv);
      const bool obs = !(y[a] != y[a]);
      const double ya = obs ? y[a] : 0.0;
      const double sd = sqrt(2.0 * fv);
      double e0 = 0.0, e1 = 0.0, e2 = 0.0;
      // the f-independent part of l(f) only shifts the ELL: skipped when the caller wants the gradients alone
      double c0 = 0.0;
      if (LIK == CVI_LIK_POISSON_EXP && want_ell)
        c0 = have_log ? poisson_exp_const_tab(ya, log_param, logfact) : poisson_exp_const(ya, lik_param);
      if (LIK == CVI_LIK_POISSON_EXP) {
        // exp(m + sd x_q) = exp(m) exp(sd x_q), and Gauss-Hermite nodes come in pairs +-x (numpy's hermgauss
        // symmetrises them exactly): one exp and one reciprocal per PAIR instead of two exps -- the quadrature
        // loop is what bounds this kernel (K = 20 fp64 exps per scalar site).  Nodes that are not an exact
        // +- pair take the plain path.
        // For sd = sqrt(2 v) <= 3 a K >= 20 point rule integrates exp(sd x) to round-off: sum_q w_q exp(sd x_q)
        // == exp(sd^2 / 4) within 4e-16 relative (tests/test_oracle_cvi.py pins this against numpy's nodes), so
        // all three sums follow from ONE exp -- the reference's own closed-form Poisson ELL
        // (expected_log_likelihoods.py:149-174): E[l] = y m + c0 - b exp(m + v / 2), E[l'] = y - b exp(m + v / 2),
        // E[l''] = -b exp(m + v / 2).  Wider marginals run the node loop.
        if (K >= 20 && sd <= 3.0) {
          const double E = exp(fma(0.25 * sd, sd, fmu[a])) * lik_param;
          e0 = fma(ya, fmu[a], c0) - E;
          e1 = ya - E;
          e2 = -E;
        } else {
        const double em = exp(fmu[a]) * lik_param;
        for (int q = 0; q < (K + 1) / 2; ++q) {
          const int r = K - 1 - q;
          const double xq = ghx[q], xr = ghx[r];
          const double ep = exp(sd * xq);
          {
            const double lam = em * ep;
            e0 = fma(ghw[q], fma(ya, fma(sd, xq, fmu[a]), c0) - lam, e0);
            e1 = fma(ghw[q], ya - lam, e1);
            e2 = fma(ghw[q], -lam, e2);
          }
          if (r != q) {
            const double er = (xr == -xq) ? fast_rcp(ep) : exp(sd * xr);
            const double lam = em * er;
            e0 = fma(ghw[r], fma(ya, fma(sd, xr, fmu[a]), c0) - lam, e0);
            e1 = fma(ghw[r], ya - lam, e1);
            e2 = fma(ghw[r], -lam, e2);
          }
        }
        }
      } else {
        for (int q = 0; q < K; ++q) {
          const double f = fma(sd, ghx[q], fmu[a]);
          double l, d1, d2;
          bernoulli_probit_terms(ya, f, l, d1, d2);
          e0 = fma(ghw[q], l, e0);
          e1 = fma(ghw[q], d1, e1);
          e2 = fma(ghw[q], d2, e2);
        }
      }
      if (obs) {
        ell += e0;
        PHYSS_UNROLL
        for (int i = 0; i < D; ++i) {
          dm[i] = fma(W[a][i], e1, dm[i]);
          PHYSS_UNROLL
          for (int j = 0; j < D; ++j) dS[i][j] = fma(0.5 * e2 * W[a][i], W[a][j], dS[i][j]);
        }
      }
    }
  }
  return ell;
}

// One natural-gradient site update (theta -> lambda, cvi_block_update, lambda -> theta).
//   prec ('NG_Precision', exponential_family_transforms.py:44-53,85-95; cvi_parameterisations.py:95-113): Vt is the site
//   PRECISION.  lambda_2 = -1/2 Vt; lambda_1 = (Vt + ng_jitter I)^-1 Y~ -- the reference's own cholesky_solve with the
//   precision's factor --; on the way back theta_2' = -2 lambda_2' (written to Vn) and
//   theta_1' = (-2 lambda_2' + ng_jitter I)^-1 lambda_1'.
template <int D>
PHYSS_HD void cvi_site_update(const double (&Yt)[D], const double (&Vt)[D][D], const double (&qm)[D],
                              const double (&qS)[D][D], const double (&dm)[D],
                              const double (&dS)[D][D], double beta, double ngj, double (&Yn)[D],
                              double (&Vn)[D][D], bool prec = false) {
  double Vinv[D][D];
  spd_inverse<D>(Vt, ngj, Vinv);                 // (V~ + ng_jitter I)^-1
  double l1[D], l2[D][D];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    double t = 0.0;
    PHYSS_UNROLL
    for (int k = 0; k < D; ++k) t = fma(Vinv[i][k], Yt[k], t);
    // grad_1 = dm - 2 dS m
    double g = dm[i];
    PHYSS_UNROLL
    for (int k = 0; k < D; ++k) g = fma(-2.0 * dS[i][k], qm[k], g);
    l1[i] = (1.0 - beta) * t + beta * g;
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) l2[i][j] = (1.0 - beta) * (-0.5 * (prec ? Vt[i][j] : Vinv[i][j])) + beta * dS[i][j];
  }
  (void)qS;
  double Pm[D][D];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) Pm[i][j] = -2.0 * l2[i][j];
  }
  spd_inverse<D>(Pm, ngj, Vn);                   // V~' = (-2 lambda_2 + ng_jitter I)^-1
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    double t = 0.0;
    PHYSS_UNROLL
    for (int k = 0; k < D; ++k) t = fma(Vn[i][k], l1[k], t);
    Yn[i] = t;
  }
  if (prec) {
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      PHYSS_UNROLL
      for (int j = 0; j < D; ++j) Vn[i][j] = Pm[i][j];
    }
  }
}

}  // namespace physs
