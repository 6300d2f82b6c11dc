/* ssm_oracle.c -- plain-C restatement of the reference's sequential Kalman filter + RTS smoother.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): used by tests/ as a fast checker for parity cases the
 * numpy oracle is too slow for, and by bench.py as the reported CPU baseline ("port").  It is never
 * linked into, imported by or called from the product.
 *
 * It follows the reference step by step with dense d x d algebra, exactly as the JAX code does
 * (paths relative to /root/reference/src/lib/stgp/):
 *   kernels/ss_utils.py:6-10, kernels/matern.py:152-177,306-329  closed-form A = expm(F dt)
 *   kernels/kernel.py:207-209                                    Q = Pinf - A Pinf A^T  (verbatim)
 *   computation/filters/kalman_filter.py:214-241                 predict
 *   computation/filters/kalman_filter.py:144-211                 masked update, jittered gain solve
 *   computation/gaussian.py:42-108                               lml with mask-to-identity
 *   computation/filters/rts_smoother.py:48-106,162-192           RTS step, jittered chol(P_pred)
 * Validated against the numpy oracle (oracle/filters.py) in tests/test_c_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 64

static void matmul(int n, int k, int m, const double* A, const double* B, double* C) {
  /* C[n,m] = A[n,k] B[k,m] */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = 0.0;
      for (int l = 0; l < k; ++l) acc += A[i * k + l] * B[l * m + j];
      C[i * m + j] = acc;
    }
}

static void matmul_nt(int n, int k, int m, const double* A, const double* B, double* C) {
  /* C[n,m] = A[n,k] B[m,k]^T */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double acc = 0.0;
      for (int l = 0; l < k; ++l) acc += A[i * k + l] * B[j * k + l];
      C[i * m + j] = acc;
    }
}

/* lower Cholesky; a non-PD matrix yields NaN (sqrt of a negative), like jnp.linalg.cholesky */
static void cholesky(int n, const double* A, double* L) {
  memset(L, 0, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    double ljj = sqrt(s);
    L[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double t = A[i * n + j];
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / ljj;
    }
  }
}

/* X[n,r] <- (L L^T)^{-1} X */
static void cho_solve(int n, int r, const double* L, double* X) {
  for (int c = 0; c < r; ++c) {
    for (int i = 0; i < n; ++i) {
      double t = X[i * r + c];
      for (int k = 0; k < i; ++k) t -= L[i * n + k] * X[k * r + c];
      X[i * r + c] = t / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double t = X[i * r + c];
      for (int k = i + 1; k < n; ++k) t -= L[k * n + i] * X[k * r + c];
      X[i * r + c] = t / L[i * n + i];
    }
  }
}

/* closed-form expm(F dt) for a Matern-(s-1/2) block */
static void matern_expm(int s, double lam, double dt, double* A) {
  double x = dt * lam, e = exp(-x);
  if (s == 1) {
    A[0] = e;
  } else if (s == 2) { /* ss_utils.py:6-10 */
    A[0] = e * (dt * lam + 1.0); A[1] = e * dt;
    A[2] = e * (dt * -(lam * lam)); A[3] = e * (dt * -lam + 1.0);
  } else if (s == 3) { /* matern.py:152-177 */
    double M[9] = {lam * (0.5 * x + 1.0), x + 1.0, 0.5 * dt,
                   -0.5 * x * lam * lam, lam * (1.0 - x), 1.0 - 0.5 * x,
                   lam * lam * lam * (0.5 * x - 1.0), lam * lam * (x - 3.0), lam * (0.5 * x - 2.0)};
    for (int i = 0; i < 9; ++i) A[i] = e * (dt * M[i] + ((i % 4 == 0) ? 1.0 : 0.0));
  } else { /* matern.py:306-329 */
    double x2 = x * x, l2 = lam * lam, l3 = l2 * lam;
    double M[16] = {lam * (1.0 + 0.5 * x + x2 / 6.0), 1.0 + x + 0.5 * x2, 0.5 * dt * (1.0 + x), dt * dt / 6.0,
                    -x2 * l2 / 6.0, lam * (1.0 + 0.5 * x - 0.5 * x2), 1.0 + x - 0.5 * x2, dt * (0.5 - x / 6.0),
                    l3 * x * (x / 6.0 - 0.5), x * l2 * (0.5 * x - 2.0), lam * (1.0 - 2.5 * x + 0.5 * x2), 1.0 - x + x2 / 6.0,
                    l2 * l2 * (x - 1.0 - x2 / 6.0), l3 * (3.5 * x - 4.0 - 0.5 * x2), l2 * (4.0 * x - 6.0 - 0.5 * x2), lam * (1.5 * x - 3.0 - x2 / 6.0)};
    for (int i = 0; i < 16; ++i) A[i] = e * (dt * M[i] + ((i % 5 == 0) ? 1.0 : 0.0));
  }
}

static void build_A(int d, int s, int nblk, const double* lam, double dt, double* A) {
  double blk[16];
  memset(A, 0, sizeof(double) * d * d);
  for (int b = 0; b < nblk; ++b) {
    matern_expm(s, lam[b], dt, blk);
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) A[(b * s + i) * d + b * s + j] = blk[i * s + j];
  }
}

typedef struct {
  double A[MAXD * MAXD], Q[MAXD * MAXD], T1[MAXD * MAXD], T2[MAXD * MAXD], T3[MAXD * MAXD];
  double L[MAXD * MAXD], S[MAXD * MAXD], Sj[MAXD * MAXD], HP[MAXD * MAXD], K[MAXD * MAXD];
  double v[MAXD], mu[MAXD], w[MAXD], mp[MAXD];
} work_t;

/* one series: filter (always) and, if ms != NULL, smoother.  Returns lml. */
static double one_series(int64_t T, int d, int m, int s, int nblk, const double* lam, const double* Pinf,
                         const double* H, const double* dt_f, const double* dt_s, const double* Y,
                         const double* R, int64_t R_ts, double jitter, int full_state, int mo,
                         double* mf, double* Pf, double* ms_out, double* Ps_out, work_t* w) {
  const double LOG2PI = 1.8378770664093454835606594728112;
  double m_[MAXD], P[MAXD * MAXD];
  memset(m_, 0, sizeof(double) * d);
  memcpy(P, Pinf, sizeof(double) * d * d);
  double lml = 0.0;
  for (int64_t k = 0; k < T; ++k) {
    build_A(d, s, nblk, lam, dt_f[k], w->A);
    matmul(d, d, d, w->A, Pinf, w->T1);
    matmul_nt(d, d, d, w->T1, w->A, w->T2);
    for (int i = 0; i < d * d; ++i) w->Q[i] = Pinf[i] - w->T2[i];
    /* predict */
    matmul(d, d, 1, w->A, m_, w->mp);
    matmul(d, d, d, w->A, P, w->T1);
    matmul_nt(d, d, d, w->T1, w->A, P);
    for (int i = 0; i < d * d; ++i) P[i] += w->Q[i];
    /* update */
    const double* y = Y + k * m;
    const double* Rk = R + k * R_ts;
    int obs[MAXD], n_missing = 0;
    for (int a = 0; a < m; ++a) { obs[a] = !(y[a] != y[a]); n_missing += !obs[a]; }
    matmul(m, d, d, H, P, w->HP);
    for (int a = 0; a < m; ++a) if (!obs[a]) for (int j = 0; j < d; ++j) w->HP[a * d + j] = 0.0;
    matmul(m, d, 1, H, w->mp, w->mu);
    for (int a = 0; a < m; ++a) { if (!obs[a]) w->mu[a] = 0.0; w->v[a] = (obs[a] ? y[a] : 0.0) - w->mu[a]; }
    matmul_nt(m, d, m, w->HP, H, w->S);
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b) {
        double var = obs[b] ? w->S[a * m + b] : 0.0;
        w->S[a * m + b] = var + Rk[a * m + b];
        w->Sj[a * m + b] = w->S[a * m + b] + (a == b ? jitter : 0.0);
      }
    cholesky(m, w->Sj, w->L);
    memcpy(w->K, w->HP, sizeof(double) * m * d); /* K^T [m,d] */
    cho_solve(m, d, w->L, w->K);
    for (int i = 0; i < d; ++i) {
      double acc = w->mp[i];
      for (int a = 0; a < m; ++a) acc += w->K[a * d + i] * w->v[a];
      m_[i] = acc;
    }
    /* P -= K S K^T: T1[d,m] = K S */
    for (int i = 0; i < d; ++i)
      for (int b = 0; b < m; ++b) {
        double acc = 0.0;
        for (int a = 0; a < m; ++a) acc += w->K[a * d + i] * w->S[a * m + b];
        w->T1[i * m + b] = acc;
      }
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) {
        double acc = 0.0;
        for (int b = 0; b < m; ++b) acc += w->T1[i * m + b] * w->K[b * d + j];
        P[i * d + j] -= acc;
      }
    /* lml */
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b)
        w->Sj[a * m + b] = (obs[a] && obs[b]) ? w->S[a * m + b] : (a == b ? 1.0 : 0.0);
    cholesky(m, w->Sj, w->L);
    double logdet = 0.0;
    for (int a = 0; a < m; ++a) logdet += log(w->L[a * m + a] * w->L[a * m + a]);
    memcpy(w->w, w->v, sizeof(double) * m);
    cho_solve(m, 1, w->L, w->w);
    double mahal = 0.0;
    for (int a = 0; a < m; ++a) mahal += w->v[a] * w->w[a];
    lml += -0.5 * m * LOG2PI - 0.5 * logdet - 0.5 * mahal + 0.5 * n_missing * LOG2PI;
    memcpy(mf + k * d, m_, sizeof(double) * d);
    memcpy(Pf + k * d * d, P, sizeof(double) * d * d);
  }
  if (!ms_out) return lml;
  /* smoother */
  const int mp = full_state ? d : mo;
  double msm[MAXD], Ps[MAXD * MAXD];
  memcpy(msm, mf + (T - 1) * d, sizeof(double) * d);
  memcpy(Ps, Pf + (T - 1) * d * d, sizeof(double) * d * d);
  for (int64_t k = T - 1; k >= 0; --k) {
    if (k < T - 1) {
      const double* mfk = mf + k * d;
      const double* Pfk = Pf + k * d * d;
      build_A(d, s, nblk, lam, dt_s[k], w->A);
      matmul(d, d, d, w->A, Pinf, w->T1);
      matmul_nt(d, d, d, w->T1, w->A, w->T2);
      for (int i = 0; i < d * d; ++i) w->Q[i] = Pinf[i] - w->T2[i];
      matmul(d, d, 1, w->A, mfk, w->mp);
      matmul(d, d, d, w->A, Pfk, w->T1);            /* A Pf */
      matmul_nt(d, d, d, w->T1, w->A, w->T2);        /* A Pf A^T */
      for (int i = 0; i < d * d; ++i) w->T2[i] += w->Q[i];  /* P_pred */
      for (int i = 0; i < d * d; ++i) w->Sj[i] = w->T2[i] + ((i % (d + 1) == 0) ? jitter : 0.0);
      cholesky(d, w->Sj, w->L);
      cho_solve(d, d, w->L, w->T1);                  /* G^T [d,d] */
      double nm[MAXD];
      for (int i = 0; i < d; ++i) {
        double acc = mfk[i];
        for (int j = 0; j < d; ++j) acc += w->T1[j * d + i] * (msm[j] - w->mp[j]);
        nm[i] = acc;
      }
      memcpy(msm, nm, sizeof(double) * d);
      for (int i = 0; i < d * d; ++i) w->T3[i] = Ps[i] - w->T2[i];
      /* W = G dP : W[i,j] = sum_k G[i,k] dP[k,j] = sum_k Gt[k,i] dP[k,j] */
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          double acc = 0.0;
          for (int l = 0; l < d; ++l) acc += w->T1[l * d + i] * w->T3[l * d + j];
          w->S[i * d + j] = acc;
        }
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          double acc = Pfk[i * d + j];
          for (int l = 0; l < d; ++l) acc += w->S[i * d + l] * w->T1[l * d + j];
          Ps[i * d + j] = acc;
        }
    }
    if (full_state) {
      memcpy(ms_out + k * d, msm, sizeof(double) * d);
      memcpy(Ps_out + k * d * d, Ps, sizeof(double) * d * d);
    } else {
      matmul(mo, d, 1, H, msm, ms_out + k * mp);
      matmul(mo, d, d, H, Ps, w->HP);
      matmul_nt(mo, d, mo, w->HP, H, Ps_out + k * mp * mp);
    }
  }
  return lml;
}

/* B independent series.  lam [B or 1, nblk], Pinf [B or 1, d, d] (stride 0 = shared), H [m, d],
 * dt_f / dt_s [T] shared grid, Y [B, T, m], R [.., m, m] with batch/time strides.
 * mf/Pf may be NULL (per-thread scratch is used); ms/Ps NULL skips the smoother.
 * Returns the number of threads used. */
int oracle_filter_smooth(int64_t B, int64_t T, int d, int m, int s, int nblk,
                         const double* lam, int64_t lam_bs, const double* Pinf, int64_t Pinf_bs,
                         const double* H, const double* dt_f, const double* dt_s,
                         const double* Y, const double* R, int64_t R_bs, int64_t R_ts,
                         double jitter, int full_state,
                         double* mf, double* Pf, double* lml, double* ms, double* Ps, int nthreads) {
  if (d > MAXD || m > MAXD) return -1;
  const int mp = full_state ? d : m;
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#endif
#pragma omp parallel
  {
    work_t* w = (work_t*)malloc(sizeof(work_t));
    double* smf = mf ? NULL : (double*)malloc(sizeof(double) * T * d);
    double* sPf = Pf ? NULL : (double*)malloc(sizeof(double) * T * d * d);
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
      double* mfb = mf ? mf + b * T * d : smf;
      double* Pfb = Pf ? Pf + b * T * d * d : sPf;
      lml[b] = one_series(T, d, m, s, nblk, lam + b * lam_bs, Pinf + b * Pinf_bs, H, dt_f, dt_s,
                          Y + b * T * m, R + b * R_bs, R_ts, jitter, full_state, m, mfb, Pfb,
                          ms ? ms + b * T * mp : NULL, Ps ? Ps + b * T * mp * mp : NULL, w);
    }
    free(w); free(smf); free(sPf);
  }
  return used;
}
