// physs_core.cuh -- register-resident per-series state-space algebra (fp64).
//
// Everything here is PHYSS_HD (__host__ __device__) and templated on compile-time sizes so that
//  (a) on the GPU every small matrix lives in registers, fully unrolled (one thread = one series);
//  (b) the identical source compiles with g++ for the CPU-side logic check in tests/ (there is no
//      GPU in the build container) -- that host build is a TEST of this file, not a product path.
//
// Reference semantics being reproduced (paths relative to /root/reference/src/lib/stgp/):
//   computation/filters/kalman_filter.py:144-241   predict + masked update + lml
//   computation/filters/rts_smoother.py:48-106     RTS step with jittered chol(P_pred)
//   computation/linalg.py:12-33                    solve() = chol(S + jitter I) solve
//   computation/gaussian.py:42-108                 log N with mask-to-identity, un-jittered chol
//   kernels/ss_utils.py:6-10, kernels/matern.py:152-177,306-329   closed-form expm(F dt)
//   kernels/kernel.py:207-209                      Q = Pinf - A Pinf A^T
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PHYSS_HD __host__ __device__ __forceinline__
#define PHYSS_UNROLL _Pragma("unroll")
#else
#define PHYSS_HD inline
#define PHYSS_UNROLL
#endif

namespace physs {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kLn2 = 0.69314718055994530941723212145818;

// Branch-free reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~2^-20)
// refined by three Newton steps to <= 2 ulp.  Unlike `1.0 / x` and `sqrt(x)` they compile to
// straight-line code (no slow-path call), which keeps each time step one schedulable basic block.
// x < 0 -> NaN, as the reference's Cholesky produces for a non-PD matrix.
PHYSS_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  // hardware seed (MUFU.RCP64H: ~2^-20 relative) + two Newton steps: 2^-40, then rounding-limited (a third step, kept
  // until the end of round 2, changed nothing but the length of the dependency chain)
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
#else
  return 1.0 / x;
#endif
}

PHYSS_HD double fast_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = fma(y, fma(-hx * y, y, 0.5), y);
  y = fma(y, fma(-hx * y, y, 0.5), y);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}

// Running log-determinant / Mahalanobis accumulator for the marginal likelihood: sum_k log|S_k| is
// carried as a (mantissa product, binary exponent) pair so that no log() sits in the time loop.
struct LmlAcc {
  double prod = 1.0;   // product of determinants, renormalised into [0.5, 1)
  long long expo = 0;  // accumulated binary exponent
  double quad = 0.0;   // sum of Mahalanobis terms
  long long nobs = 0;  // number of observed scalars
  PHYSS_HD void add(double det, double mahal, int n_obs) {
    prod *= det;
    int e;
    prod = frexp(prod, &e);
    expo += e;
    quad += mahal;
    nobs += n_obs;
  }
  PHYSS_HD double value() const {
    return -0.5 * ((double)nobs * kLog2Pi + (log(prod) + (double)expo * kLn2) + quad);
  }
};

// ---------------------------------------------------------------------------------------------
// Closed-form A = expm(F dt) for Matern-(S-1/2) state-space blocks, S = 1..4.
// lam = sqrt(2 nu) / lengthscale.  S=2: ss_utils.py:6-10; S=3: matern.py:152-177; S=4: matern.py:306-329.
// ---------------------------------------------------------------------------------------------
// Forward-mode pair (value, derivative) for d/dlam of the closed forms (used by the adjoint kernel, physs_vjp.cu)
struct Dual {
  double v, d;
};
PHYSS_HD Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
PHYSS_HD Dual operator+(Dual a, double b) { return {a.v + b, a.d}; }
PHYSS_HD Dual operator+(double a, Dual b) { return {a + b.v, b.d}; }
PHYSS_HD Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
PHYSS_HD Dual operator-(Dual a, double b) { return {a.v - b, a.d}; }
PHYSS_HD Dual operator-(double a, Dual b) { return {a - b.v, -b.d}; }
PHYSS_HD Dual operator-(Dual a) { return {-a.v, -a.d}; }
PHYSS_HD Dual operator*(Dual a, Dual b) { return {a.v * b.v, fma(a.v, b.d, a.d * b.v)}; }
PHYSS_HD Dual operator*(Dual a, double b) { return {a.v * b, a.d * b}; }
PHYSS_HD Dual operator*(double a, Dual b) { return {a * b.v, a * b.d}; }
PHYSS_HD Dual operator/(Dual a, double b) { return {a.v / b, a.d / b}; }
PHYSS_HD double texp(double x) { return exp(x); }
PHYSS_HD Dual texp(Dual x) { const double e = exp(x.v); return {e, e * x.d}; }

template <int S>
struct MaternExpm;

template <>
struct MaternExpm<1> {
  static PHYSS_HD void eval(double lam, double dt, double (&A)[1][1]) { evalT<double>(lam, dt, A); }
  template <class T>
  static PHYSS_HD void evalT(T lam, double dt, T (&A)[1][1]) { A[0][0] = texp(-(lam * dt)); }
};

template <>
struct MaternExpm<2> {
  static PHYSS_HD void eval(double lam, double dt, double (&A)[2][2]) { evalT<double>(lam, dt, A); }
  template <class T>
  static PHYSS_HD void evalT(T lam, double dt, T (&A)[2][2]) {
    const T e = texp(-(dt * lam));
    A[0][0] = e * (dt * lam + 1.0);
    A[0][1] = e * dt;
    A[1][0] = e * (dt * (-lam * lam));
    A[1][1] = e * (dt * (-lam) + 1.0);
  }
};

// Size-2 block of a DISC_MATERN stack on the lane-group paths: lam >= 0 is a Matern-3/2 block; a lam with the SIGN BIT
// set (including -0.0) selects a harmonic-oscillator block with angular frequency -lam, F = [[0, -w], [w, 0]],
// expm(F dt) = rotation by w dt -- the j-th component of the reference's periodic kernels (kernels/periodic.py:
// 213-253, F = kron(diag(0..J), [[0, -w0], [w0, 0]]), which the reference pushes through the generic
// jax.scipy.linalg.expm).  Pinf of such a block is q_j^2 I, so Q_k = Pinf - A Pinf A^T vanishes to round-off.
PHYSS_HD void block2_expm(double lam, double dt, double (&A)[2][2]) {
  if (signbit(lam)) {
    double sn, cs;
    sincos(-lam * dt, &sn, &cs);
    A[0][0] = cs; A[0][1] = -sn;
    A[1][0] = sn; A[1][1] = cs;
  } else {
    MaternExpm<2>::eval(lam, dt, A);
  }
}

template <>
struct MaternExpm<3> {
  static PHYSS_HD void eval(double lam, double dt, double (&A)[3][3]) { evalT<double>(lam, dt, A); }
  template <class T>
  static PHYSS_HD void evalT(T lam, double dt, T (&A)[3][3]) {
    const T x = dt * lam;  // dtlam
    const T e = texp(-x);
    const T l2 = lam * lam;
    A[0][0] = e * (dt * (lam * (0.5 * x + 1.0)) + 1.0);
    A[0][1] = e * (dt * (x + 1.0));
    A[0][2] = e * (dt * (0.5 * dt));
    A[1][0] = e * (dt * (-0.5 * x * l2));
    A[1][1] = e * (dt * (lam * (1.0 - x)) + 1.0);
    A[1][2] = e * (dt * (1.0 - 0.5 * x));
    A[2][0] = e * (dt * (l2 * lam * (0.5 * x - 1.0)));
    A[2][1] = e * (dt * (l2 * (x - 3.0)));
    A[2][2] = e * (dt * (lam * (0.5 * x - 2.0)) + 1.0);
  }
};

template <>
struct MaternExpm<4> {
  static PHYSS_HD void eval(double lam, double dt, double (&A)[4][4]) { evalT<double>(lam, dt, A); }
  template <class T>
  static PHYSS_HD void evalT(T lam, double dt, T (&A)[4][4]) {
    const T x = dt * lam;
    const T x2 = x * x;
    const T e = texp(-x);
    const T l2 = lam * lam;
    const T l3 = l2 * lam;
    // 1/6 as a multiplication: a double division is a ~25-instruction sequence with a slow-path BRANCH that
    // splits the time step into several basic blocks (no scheduling across them); the products differ from the
    // reference's quotients by <= 1 ulp
    constexpr double kSixth = 1.0 / 6.0;
    A[0][0] = e * (dt * (lam * (1.0 + 0.5 * x + x2 * kSixth)) + 1.0);
    A[0][1] = e * (dt * (1.0 + x + 0.5 * x2));
    A[0][2] = e * (dt * (0.5 * dt * (1.0 + x)));
    A[0][3] = e * (dt * (dt * dt * kSixth));
    A[1][0] = e * (dt * (-x2 * l2 * kSixth));
    A[1][1] = e * (dt * (lam * (1.0 + 0.5 * x - 0.5 * x2)) + 1.0);
    A[1][2] = e * (dt * (1.0 + x - 0.5 * x2));
    A[1][3] = e * (dt * (dt * (0.5 - x * kSixth)));
    A[2][0] = e * (dt * (l3 * x * (x * kSixth - 0.5)));
    A[2][1] = e * (dt * (x * l2 * (0.5 * x - 2.0)));
    A[2][2] = e * (dt * (lam * (1.0 - 2.5 * x + 0.5 * x2)) + 1.0);
    A[2][3] = e * (dt * (1.0 - x + x2 * kSixth));
    A[3][0] = e * (dt * (l2 * l2 * (x - 1.0 - x2 * kSixth)));
    A[3][1] = e * (dt * (l3 * (3.5 * x - 4.0 - 0.5 * x2)));
    A[3][2] = e * (dt * (l2 * (4.0 * x - 6.0 - 0.5 * x2)));
    A[3][3] = e * (dt * (lam * (1.5 * x - 3.0 - x2 * kSixth)) + 1.0);
  }
};

// ---------------------------------------------------------------------------------------------
// Integrated Wiener process of order q = S - 1 (WienerVelocity, kernels/wiener.py:60-149): F = shift matrix,
//   A[i][j] = dt^(j-i) / (j-i)!  (j >= i, else 0)                                         (:105-123)
//   Q[i][j] = var * dt^(2q+1-i-j) / ((2q+1-i-j) (q-i)! (q-j)!)                           (:125-149)
// Not stationary: there is no P_inf with Q = P_inf - A P_inf A^T, so the predict step takes (A, Q) explicitly.
// ---------------------------------------------------------------------------------------------
template <int S>
struct IwpDisc {
  static PHYSS_HD void eval(double var, double dt, double (&A)[S][S], double (&Q)[S][S]) {
    constexpr int q = S - 1;
    constexpr double fact[8] = {1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 720.0, 5040.0};
    double pw[2 * q + 2];                       // dt^0 .. dt^(2q+1)
    pw[0] = 1.0;
    PHYSS_UNROLL
    for (int k = 1; k < 2 * q + 2; ++k) pw[k] = pw[k - 1] * dt;
    PHYSS_UNROLL
    for (int i = 0; i < S; ++i) {
      PHYSS_UNROLL
      for (int j = 0; j < S; ++j) {
        A[i][j] = (j >= i) ? pw[(j >= i) ? j - i : 0] / fact[(j >= i) ? j - i : 0] : 0.0;
        const int e = 2 * q + 1 - i - j;
        Q[i][j] = var * pw[e] / ((double)e * fact[q - i] * fact[q - j]);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Block-diagonal transition: D = NB * S, block b is an S x S dense matrix.  S == D is a plain dense
// transition (also used when A_k is supplied by the caller).
// ---------------------------------------------------------------------------------------------
template <int D, int S>
struct Trans {
  static constexpr int NB = D / S;
  double a[NB][S][S];

  // y = A x
  PHYSS_HD void mulv(const double (&x)[D], double (&y)[D]) const {
    PHYSS_UNROLL
    for (int b = 0; b < NB; ++b) {
      PHYSS_UNROLL
      for (int i = 0; i < S; ++i) {
        double acc = 0.0;
        PHYSS_UNROLL
        for (int k = 0; k < S; ++k) acc = fma(a[b][i][k], x[b * S + k], acc);
        y[b * S + i] = acc;
      }
    }
  }
  // C = A X   (X, C dense D x D)
  PHYSS_HD void mulL(const double (&X)[D][D], double (&C)[D][D]) const {
    PHYSS_UNROLL
    for (int b = 0; b < NB; ++b) {
      PHYSS_UNROLL
      for (int i = 0; i < S; ++i) {
        PHYSS_UNROLL
        for (int j = 0; j < D; ++j) {
          double acc = 0.0;
          PHYSS_UNROLL
          for (int k = 0; k < S; ++k) acc = fma(a[b][i][k], X[b * S + k][j], acc);
          C[b * S + i][j] = acc;
        }
      }
    }
  }
  // Out = Add + C A^T, only the upper triangle is computed and mirrored (result is symmetric by
  // construction: C = A X with X symmetric).
  PHYSS_HD void mulRT_sym_add(const double (&C)[D][D], const double (&Add)[D][D],
                              double (&Out)[D][D]) const {
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      PHYSS_UNROLL
      for (int j = i; j < D; ++j) {
        const int b = j / S, jj = j % S;
        double acc = Add[i][j];
        PHYSS_UNROLL
        for (int k = 0; k < S; ++k) acc = fma(C[i][b * S + k], a[b][jj][k], acc);
        Out[i][j] = acc;
        Out[j][i] = acc;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Small dense Cholesky / triangular solves (lower), in registers.  Non-PD input -> NaN (sqrt of a
// negative number), which then propagates exactly as jnp.linalg.cholesky's NaNs do.
// ---------------------------------------------------------------------------------------------
template <int N>
PHYSS_HD void chol_lower(const double (&A)[N][N], double (&L)[N][N], double (&rdiag)[N]) {
  PHYSS_UNROLL
  for (int j = 0; j < N; ++j) {
    double s = A[j][j];
    PHYSS_UNROLL
    for (int k = 0; k < j; ++k) s = fma(-L[j][k], L[j][k], s);
    const double r = fast_rsqrt(s);
    const double ljj = s * r;
    L[j][j] = ljj;
    rdiag[j] = r;
    PHYSS_UNROLL
    for (int i = j + 1; i < N; ++i) {
      double t = A[i][j];
      PHYSS_UNROLL
      for (int k = 0; k < j; ++k) t = fma(-L[i][k], L[j][k], t);
      L[i][j] = t * r;
    }
  }
}

// Solve (L L^T) x = b in place for one right-hand side.
template <int N>
PHYSS_HD void chol_solve_vec(const double (&L)[N][N], const double (&rdiag)[N], double (&x)[N]) {
  PHYSS_UNROLL
  for (int i = 0; i < N; ++i) {
    double t = x[i];
    PHYSS_UNROLL
    for (int k = 0; k < i; ++k) t = fma(-L[i][k], x[k], t);
    x[i] = t * rdiag[i];
  }
  PHYSS_UNROLL
  for (int i = N - 1; i >= 0; --i) {
    double t = x[i];
    PHYSS_UNROLL
    for (int k = i + 1; k < N; ++k) t = fma(-L[k][i], x[k], t);
    x[i] = t * rdiag[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Kalman update (kalman_filter.py:144-211) for a D-state, M-observation step.
//   HID: H is the identity (full-state pseudo-observations, M == D) -- skips the H products.
// y may contain NaN (= missing).  Returns this step's log marginal likelihood term.
// ---------------------------------------------------------------------------------------------
//   mu_ext != nullptr: the predicted observation is supplied by the caller instead of H m -- the collocation
//   update of the PDE filter passes the non-linear residual g(m_) with H = dg/dx (kalman_filter.py:395-414).
template <int D, int M, bool HID>
PHYSS_HD void kf_update(double (&m)[D], double (&P)[D][D], const double (&H)[M][D],
                        const double (&R)[M][M], const double (&y)[M], double jitter,
                        double& det_out, double& mahal_out, int& nobs_out, const double* mu_ext = nullptr) {
  // mask
  bool obs[M];
  int n_missing = 0;
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) {
    obs[a] = !(y[a] != y[a]);
    n_missing += obs[a] ? 0 : 1;
  }
  // HP = M H P_  (rows of missing observations zeroed), innovation
  double HP[M][D];
  double v[M];
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) {
    double mu = 0.0;
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) {
      double acc;
      if (HID) {
        acc = P[a][j];
      } else {
        acc = 0.0;
        PHYSS_UNROLL
        for (int k = 0; k < D; ++k) acc = fma(H[a][k], P[k][j], acc);
      }
      HP[a][j] = obs[a] ? acc : 0.0;
    }
    if (mu_ext) {
      mu = mu_ext[a];
    } else if (HID) {
      mu = m[a];
    } else {
      PHYSS_UNROLL
      for (int k = 0; k < D; ++k) mu = fma(H[a][k], m[k], mu);
    }
    v[a] = obs[a] ? (y[a] - mu) : 0.0;
  }
  // S = M H P_ H^T M + R     (R is not masked)
  double S[M][M];
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) {
    PHYSS_UNROLL
    for (int b = a; b < M; ++b) {
      double acc;
      if (HID) {
        acc = HP[a][b];
      } else {
        acc = 0.0;
        PHYSS_UNROLL
        for (int k = 0; k < D; ++k) acc = fma(HP[a][k], H[b][k], acc);
      }
      acc = obs[b] ? acc : 0.0;
      S[a][b] = acc + R[a][b];
      S[b][a] = acc + R[b][a];
    }
  }
  if (M == 1) {
    // scalar fast path
    const double rSj = fast_rcp(S[0][0] + jitter);
    double K[D];
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) K[i] = HP[0][i] * rSj;
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) m[i] = fma(K[i], v[0], m[i]);
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      const double ks = K[i] * S[0][0];
      PHYSS_UNROLL
      for (int j = i; j < D; ++j) {
        const double pij = fma(-ks, K[j], P[i][j]);
        P[i][j] = pij;
        P[j][i] = pij;
      }
    }
    // masked step: mask_to_identity(S) = 1, v = 0  ->  lml_k = 0
    const double Sl = obs[0] ? S[0][0] : 1.0;
    // a non-positive S is a failed Cholesky in the reference (gaussian.py:56-57): NaN, not a sign that a later
    // step's negative S could cancel in the running determinant product
    det_out = (Sl > 0.0) ? Sl : nan("");
    mahal_out = v[0] * v[0] * fast_rcp(Sl);
    nobs_out = obs[0] ? 1 : 0;
  } else {
    // K^T = (S + jitter I)^{-1} (M H P_)
    double Sj[M][M];
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) {
      PHYSS_UNROLL
      for (int b = 0; b < M; ++b) Sj[a][b] = S[a][b] + (a == b ? jitter : 0.0);
    }
    double L[M][M], rd[M];
    chol_lower<M>(Sj, L, rd);
    double Kt[M][D];  // Kt[a][i] = K[i][a]
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      double x[M];
      PHYSS_UNROLL
      for (int a = 0; a < M; ++a) x[a] = HP[a][i];
      chol_solve_vec<M>(L, rd, x);
      PHYSS_UNROLL
      for (int a = 0; a < M; ++a) Kt[a][i] = x[a];
    }
    // m += K v
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      double acc = m[i];
      PHYSS_UNROLL
      for (int a = 0; a < M; ++a) acc = fma(Kt[a][i], v[a], acc);
      m[i] = acc;
    }
    // P -= K S K^T   (un-jittered S)
    double KS[D][M];
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      PHYSS_UNROLL
      for (int b = 0; b < M; ++b) {
        double acc = 0.0;
        PHYSS_UNROLL
        for (int a = 0; a < M; ++a) acc = fma(Kt[a][i], S[a][b], acc);
        KS[i][b] = acc;
      }
    }
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      PHYSS_UNROLL
      for (int j = i; j < D; ++j) {
        double acc = P[i][j];
        PHYSS_UNROLL
        for (int b = 0; b < M; ++b) acc = fma(-KS[i][b], Kt[b][j], acc);
        P[i][j] = acc;
        P[j][i] = acc;
      }
    }
    // lml: Cholesky of the UN-jittered S with missing rows/cols replaced by identity.
    double Sm[M][M];
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) {
      PHYSS_UNROLL
      for (int b = 0; b < M; ++b) {
        const bool keep = obs[a] && obs[b];
        Sm[a][b] = keep ? S[a][b] : (a == b ? 1.0 : 0.0);
      }
    }
    double L2[M][M], rd2[M];
    chol_lower<M>(Sm, L2, rd2);
    double det = 1.0;
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) det *= L2[a][a] * L2[a][a];
    double w[M];
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) w[a] = v[a];
    chol_solve_vec<M>(L2, rd2, w);
    double mahal = 0.0;
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) mahal = fma(v[a], w[a], mahal);
    det_out = det;
    mahal_out = mahal;
    nobs_out = M - n_missing;
  }
}

// ---------------------------------------------------------------------------------------------
// One step of a parallel-in-time chunk summary (physs_pscan.cu): the accumulated scan element
// (A, b, C, J, eta) of parallel_kalman_filter.py:143-220 combined with the single-step element of this
// observation.  Algebraically that is the masked Kalman update above applied to (b, C) -- same jittered
// gain, same C - K S K^T -- plus   A <- (I - K H) A,  J += (HA)^T (S+jit)^-1 (HA),  eta += (HA)^T (S+jit)^-1 v
// with HA = M H A (rows of missing observations zeroed).  On entry A, b, C are the PREDICTED quantities.
// ---------------------------------------------------------------------------------------------
template <int D, int M, bool HID>
PHYSS_HD void kf_update_summary(double (&b)[D], double (&C)[D][D], double (&A)[D][D], double (&J)[D][D],
                                double (&eta)[D], const double (&H)[M][D], const double (&R)[M][M],
                                const double (&y)[M], double jitter) {
  bool obs[M];
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) obs[a] = !(y[a] != y[a]);
  double HP[M][D], HA[M][D], v[M];
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) {
    double mu = 0.0;
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) {
      double hc, ha;
      if (HID) {
        hc = C[a][j];
        ha = A[a][j];
      } else {
        hc = 0.0;
        ha = 0.0;
        PHYSS_UNROLL
        for (int k = 0; k < D; ++k) {
          hc = fma(H[a][k], C[k][j], hc);
          ha = fma(H[a][k], A[k][j], ha);
        }
      }
      HP[a][j] = obs[a] ? hc : 0.0;
      HA[a][j] = obs[a] ? ha : 0.0;
    }
    if (HID) {
      mu = b[a];
    } else {
      PHYSS_UNROLL
      for (int k = 0; k < D; ++k) mu = fma(H[a][k], b[k], mu);
    }
    v[a] = obs[a] ? (y[a] - mu) : 0.0;
  }
  double S[M][M], Sj[M][M];
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) {
    PHYSS_UNROLL
    for (int c = a; c < M; ++c) {
      double acc;
      if (HID) {
        acc = HP[a][c];
      } else {
        acc = 0.0;
        PHYSS_UNROLL
        for (int k = 0; k < D; ++k) acc = fma(HP[a][k], H[c][k], acc);
      }
      acc = obs[c] ? acc : 0.0;
      S[a][c] = acc + R[a][c];
      S[c][a] = acc + R[c][a];
    }
  }
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) {
    PHYSS_UNROLL
    for (int c = 0; c < M; ++c) Sj[a][c] = S[a][c] + (a == c ? jitter : 0.0);
  }
  double L[M][M], rd[M];
  chol_lower<M>(Sj, L, rd);
  double Kt[M][D], Z[M][D], w[M];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    double x[M], z[M];
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) { x[a] = HP[a][i]; z[a] = HA[a][i]; }
    chol_solve_vec<M>(L, rd, x);
    chol_solve_vec<M>(L, rd, z);
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) { Kt[a][i] = x[a]; Z[a][i] = z[a]; }
  }
  PHYSS_UNROLL
  for (int a = 0; a < M; ++a) w[a] = v[a];
  chol_solve_vec<M>(L, rd, w);
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    double accb = b[i], acce = eta[i];
    PHYSS_UNROLL
    for (int a = 0; a < M; ++a) {
      accb = fma(Kt[a][i], v[a], accb);
      acce = fma(HA[a][i], w[a], acce);
    }
    b[i] = accb;
    eta[i] = acce;
  }
  double KS[D][M];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int c = 0; c < M; ++c) {
      double acc = 0.0;
      PHYSS_UNROLL
      for (int a = 0; a < M; ++a) acc = fma(Kt[a][i], S[a][c], acc);
      KS[i][c] = acc;
    }
  }
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) {
      double av = A[i][j];
      PHYSS_UNROLL
      for (int a = 0; a < M; ++a) av = fma(-Kt[a][i], HA[a][j], av);
      A[i][j] = av;
    }
    PHYSS_UNROLL
    for (int j = i; j < D; ++j) {
      double cv = C[i][j], jv = J[i][j];
      PHYSS_UNROLL
      for (int a = 0; a < M; ++a) {
        cv = fma(-KS[i][a], Kt[a][j], cv);
        jv = fma(HA[a][i], Z[a][j], jv);
      }
      C[i][j] = cv; C[j][i] = cv;
      J[i][j] = jv; J[j][i] = jv;
    }
  }
}

// log N term of one step from its (det, mahal, nobs) triple (only used when per-step values are asked for)
PHYSS_HD double lml_term(double det, double mahal, int nobs) {
  return -0.5 * ((double)nobs * kLog2Pi + log(det) + mahal);
}

// ---------------------------------------------------------------------------------------------
// Predict step.  Two equivalent forms:
//   given-Q   : P_ = A P A^T + Q                       (kalman_filter.py:234-235 verbatim)
//   stationary: P_ = Pinf + A (P - Pinf) A^T           (= A P A^T + (Pinf - A Pinf A^T), kernel.py:207-209,
//               re-associated so that Q_k never has to be formed; differs by O(eps |Pinf|))
// ---------------------------------------------------------------------------------------------
template <int D, int S>
PHYSS_HD void kf_predict_givenQ(const Trans<D, S>& A, const double (&Q)[D][D], double (&m)[D],
                                double (&P)[D][D]) {
  double m_[D];
  A.mulv(m, m_);
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) m[i] = m_[i];
  double C[D][D];
  A.mulL(P, C);
  A.mulRT_sym_add(C, Q, P);
}

template <int D, int S>
PHYSS_HD void kf_predict_stationary(const Trans<D, S>& A, const double (&Pinf)[D][D],
                                    double (&m)[D], double (&P)[D][D]) {
  double m_[D];
  A.mulv(m, m_);
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) m[i] = m_[i];
  double dP[D][D];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) dP[i][j] = P[i][j] - Pinf[i][j];
  }
  double C[D][D];
  A.mulL(dP, C);
  A.mulRT_sym_add(C, Pinf, P);
}

// ---------------------------------------------------------------------------------------------
// RTS step (rts_smoother.py:48-106).  On entry (ms, Ps) is the smoothed state at k+1, (mf, Pf) the
// filtered state at k; on exit (ms, Ps) is the smoothed state at k.
//   Qadd: the matrix added to A Pf A^T, i.e. Q_k (given) -- for the stationary form pass
//         stationary = true and Qadd = Pinf, which evaluates Pinf + (A Pf - A Pinf) A^T.
// ---------------------------------------------------------------------------------------------
// The step is split in two halves so that callers can software-pipeline it (physs_seq_impl.cuh):
//   rts_front: everything that depends on the FILTERED moments only -- m_pred, P_pred, the gain G;
//   rts_back : the two-line recursion that consumes the smoothed state of step k+1.
// Successive fronts are independent of each other; only the backs form the sequential chain.
template <int D, int S>
PHYSS_HD void rts_front(const Trans<D, S>& A, const double (&Qadd)[D][D], bool stationary,
                        const double (&mf)[D], const double (&Pf)[D][D], double jitter,
                        double (&mp)[D], double (&Pp)[D][D], double (&G)[D][D]) {
  A.mulv(mf, mp);
  double C[D][D];  // A Pf
  A.mulL(Pf, C);
  if (stationary) {
    double E[D][D];
    A.mulL(Qadd, E);  // A Pinf
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      PHYSS_UNROLL
      for (int j = 0; j < D; ++j) E[i][j] = C[i][j] - E[i][j];
    }
    A.mulRT_sym_add(E, Qadd, Pp);
  } else {
    A.mulRT_sym_add(C, Qadd, Pp);
  }
  // chol(Pp + jitter I)
  double Pj[D][D];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) Pj[i][j] = Pp[i][j] + (i == j ? jitter : 0.0);
  }
  double L[D][D], rd[D];
  chol_lower<D>(Pj, L, rd);
  // G^T = (Pp + jit)^{-1} (A Pf)  -> column j of X solves for row j of G
  PHYSS_UNROLL
  for (int j = 0; j < D; ++j) {
    double x[D];
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) x[i] = C[i][j];
    chol_solve_vec<D>(L, rd, x);
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) G[j][i] = x[i];
  }
}

template <int D>
PHYSS_HD void rts_back(const double (&mf)[D], const double (&Pf)[D][D], const double (&mp)[D],
                       const double (&Pp)[D][D], const double (&G)[D][D], double (&ms)[D], double (&Ps)[D][D]) {
  // m = mf + G (ms - mp)
  double dm[D];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) dm[i] = ms[i] - mp[i];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    double acc = mf[i];
    PHYSS_UNROLL
    for (int k = 0; k < D; ++k) acc = fma(G[i][k], dm[k], acc);
    ms[i] = acc;
  }
  // P = Pf + G (Ps - Pp) G^T
  double dP[D][D];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) dP[i][j] = Ps[i][j] - Pp[i][j];
  }
  double W[D][D];
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int j = 0; j < D; ++j) {
      double acc = 0.0;
      PHYSS_UNROLL
      for (int k = 0; k < D; ++k) acc = fma(G[i][k], dP[k][j], acc);
      W[i][j] = acc;
    }
  }
  PHYSS_UNROLL
  for (int i = 0; i < D; ++i) {
    PHYSS_UNROLL
    for (int j = i; j < D; ++j) {
      double acc = Pf[i][j];
      PHYSS_UNROLL
      for (int k = 0; k < D; ++k) acc = fma(W[i][k], G[j][k], acc);
      Ps[i][j] = acc;
      Ps[j][i] = acc;
    }
  }
}

//   Eacc != nullptr (parallel-in-time chunk summary, parallel_rts_smoother.py:25-55): additionally
//   Eacc <- G Eacc, so that after folding a chunk  x_s[first] = Eacc x + ms,  P_s[first] = Eacc P Eacc^T + Ps.
template <int D, int S>
PHYSS_HD void rts_step(const Trans<D, S>& A, const double (&Qadd)[D][D], bool stationary,
                       const double (&mf)[D], const double (&Pf)[D][D], double jitter,
                       double (&ms)[D], double (&Ps)[D][D], double (*Eacc)[D] = nullptr) {
  double mp[D], Pp[D][D], G[D][D];
  rts_front<D, S>(A, Qadd, stationary, mf, Pf, jitter, mp, Pp, G);
  if (Eacc) {
    double GE[D][D];
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      PHYSS_UNROLL
      for (int j = 0; j < D; ++j) {
        double acc = 0.0;
        PHYSS_UNROLL
        for (int k = 0; k < D; ++k) acc = fma(G[i][k], Eacc[k][j], acc);
        GE[i][j] = acc;
      }
    }
    PHYSS_UNROLL
    for (int i = 0; i < D; ++i) {
      PHYSS_UNROLL
      for (int j = 0; j < D; ++j) Eacc[i][j] = GE[i][j];
    }
  }
  rts_back<D>(mf, Pf, mp, Pp, G, ms, Ps);
}

}  // namespace physs
