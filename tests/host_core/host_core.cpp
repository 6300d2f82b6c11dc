// Host build of physs_core.cuh: checks the register-level algebra used by the CUDA kernels against
// the numpy oracle on the CPU (the build container has no GPU).  TEST ONLY -- never shipped.
#include <stdint.h>
#include <string.h>

#include "../../physs_gp_b200/csrc/physs_core.cuh"

using namespace physs;

template <int D, int S, int M, bool HID, bool GIVEN>
static void run_filter(int64_t T, const double* A_, const double* Q_, const double* lam_, const double* dt,
                       const double* Pinf_, const double* m0, const double* P0, const double* H_,
                       const double* Y, const double* R_, int64_t R_ts, double jitter, double* mf,
                       double* Pf, double* lml_k, double* lml_total) {
  double m[D], P[D][D], Pinf[D][D], H[M][D], lam[D / S];
  memcpy(m, m0, sizeof(m));
  memcpy(P, P0, sizeof(P));
  if (!GIVEN) { memcpy(Pinf, Pinf_, sizeof(Pinf)); memcpy(lam, lam_, sizeof(lam)); }
  if (!HID) memcpy(H, H_, sizeof(H));
  LmlAcc acc;
  for (int64_t k = 0; k < T; ++k) {
    double y[M], R[M][M];
    memcpy(y, Y + k * M, sizeof(y));
    memcpy(R, R_ + k * R_ts, sizeof(R));
    Trans<D, S> A;
    if constexpr (GIVEN) {
      double Q[D][D];
      memcpy(A.a, A_ + k * D * D, sizeof(double) * D * D);
      memcpy(Q, Q_ + k * D * D, sizeof(Q));
      kf_predict_givenQ<D, S>(A, Q, m, P);
    } else {
      for (int b = 0; b < D / S; ++b) MaternExpm<S>::eval(lam[b], dt[k], A.a[b]);
      kf_predict_stationary<D, S>(A, Pinf, m, P);
    }
    double det, mahal;
    int nobs;
    kf_update<D, M, HID>(m, P, H, R, y, jitter, det, mahal, nobs);
    acc.add(det, mahal, nobs);
    lml_k[k] = lml_term(det, mahal, nobs);
    memcpy(mf + k * D, m, sizeof(m));
    memcpy(Pf + k * D * D, P, sizeof(P));
  }
  *lml_total = acc.value();
}

template <int D, int S, bool GIVEN>
static void run_smooth(int64_t T, const double* A_, const double* Q_, const double* lam_, const double* dt,
                       const double* Pinf_, const double* mf, const double* Pf, double jitter,
                       double* ms_out, double* Ps_out) {
  double Pinf[D][D], lam[D / S];
  if (!GIVEN) { memcpy(Pinf, Pinf_, sizeof(Pinf)); memcpy(lam, lam_, sizeof(lam)); }
  double ms[D], Ps[D][D];
  memcpy(ms, mf + (T - 1) * D, sizeof(ms));
  memcpy(Ps, Pf + (T - 1) * D * D, sizeof(Ps));
  memcpy(ms_out + (T - 1) * D, ms, sizeof(ms));
  memcpy(Ps_out + (T - 1) * D * D, Ps, sizeof(Ps));
  for (int64_t k = T - 2; k >= 0; --k) {
    double mfk[D], Pfk[D][D];
    memcpy(mfk, mf + k * D, sizeof(mfk));
    memcpy(Pfk, Pf + k * D * D, sizeof(Pfk));
    Trans<D, S> A;
    if constexpr (GIVEN) {
      double Q[D][D];
      memcpy(A.a, A_ + k * D * D, sizeof(double) * D * D);
      memcpy(Q, Q_ + k * D * D, sizeof(Q));
      rts_step<D, S>(A, Q, false, mfk, Pfk, jitter, ms, Ps);
    } else {
      for (int b = 0; b < D / S; ++b) MaternExpm<S>::eval(lam[b], dt[k], A.a[b]);
      rts_step<D, S>(A, Pinf, true, mfk, Pfk, jitter, ms, Ps);
    }
    memcpy(ms_out + k * D, ms, sizeof(ms));
    memcpy(Ps_out + k * D * D, Ps, sizeof(Ps));
  }
}

#define CASE_F(D, S, M, HID, GIVEN)                                                            \
  if (d == D && s == S && m == M && hid == HID && given == GIVEN) {                             \
    run_filter<D, S, M, HID, GIVEN>(T, A, Q, lam, dt, Pinf, m0, P0, H, Y, R, R_ts, jitter, mf, Pf, \
                                    lml_k, lml_total);                                                   \
    return 0;                                                                                   \
  }
#define CASE_S(D, S, GIVEN)                                                      \
  if (d == D && s == S && given == GIVEN) {                                      \
    run_smooth<D, S, GIVEN>(T, A, Q, lam, dt, Pinf, mf, Pf, jitter, ms, Ps);     \
    return 0;                                                                    \
  }

extern "C" int host_filter(int d, int s, int m, int hid, int given, int64_t T, const double* A,
                           const double* Q, const double* lam, const double* dt, const double* Pinf,
                           const double* m0, const double* P0, const double* H, const double* Y,
                           const double* R, int64_t R_ts, double jitter, double* mf, double* Pf,
                           double* lml_k, double* lml_total) {
  CASE_F(2, 2, 1, false, false) CASE_F(2, 2, 1, false, true) CASE_F(2, 2, 2, true, false)
  CASE_F(3, 3, 1, false, false) CASE_F(3, 3, 3, true, false) CASE_F(3, 3, 2, false, true)
  CASE_F(4, 4, 1, false, false) CASE_F(4, 2, 1, false, false) CASE_F(4, 2, 2, false, false)
  CASE_F(4, 4, 4, true, false) CASE_F(4, 4, 4, false, true) CASE_F(4, 4, 3, false, false)
  CASE_F(4, 1, 2, false, false)
  return 1;
}

extern "C" int host_smooth(int d, int s, int given, int64_t T, const double* A, const double* Q,
                           const double* lam, const double* dt, const double* Pinf, const double* mf,
                           const double* Pf, double jitter, double* ms, double* Ps) {
  CASE_S(2, 2, false) CASE_S(2, 2, true) CASE_S(3, 3, false) CASE_S(3, 3, true)
  CASE_S(4, 4, false) CASE_S(4, 2, false) CASE_S(4, 4, true) CASE_S(4, 1, false)
  return 1;
}
