// physs_api.cu -- the extern "C" surface of libphyss_b200.so (see include/physs_b200.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "physs_internal.h"

namespace physs {

static thread_local char g_err[256] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PHYSS_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return PHYSS_ERR_CUDA;
}

}  // namespace physs

namespace physs {

// PHYSS_FORCE_GRP=1 (debug / A-B timing only): route sizes both families cover to the lane-group kernels
static bool force_grp(int d) {
  static const bool on = [] { const char* e = getenv("PHYSS_FORCE_GRP"); return e && e[0] == '1'; }();
  return on && d > 4;
}

// d <= 4: one thread per series, everything in registers.  d = 8 is covered by both families:
static bool prefer_seq(int d, int m, int64_t B, int64_t nchunk, bool smoother) {
  if (d <= 4) return true;
  // d = 8 thread-per-series variant: superseded by the lane-group kernels with compile-time shapes (measured at
  // 7104 series x 10000 steps: filter 53 ms vs 32 ms, smoother 176 ms vs 72 ms); PHYSS_SEQ8=1 keeps it reachable
  // for A-B timing
  static const bool seq8 = [] { const char* e = getenv("PHYSS_SEQ8"); return e && e[0] == '1'; }();
  (void)smoother;
  return seq8 && nchunk == 0 && B >= 4096 && m <= 4;
}

int run_filter_any(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity, const SeqFilterArgs& a) {
  if (disc_mode == PHYSS_DISC_IWP && !seq_supported(d, m, disc_mode, nblk))
    return set_error(PHYSS_ERR_UNSUPPORTED, "DISC_IWP: one block of state dim 2..4 (use DISC_GIVEN otherwise)");
  if (!force_grp(d) && prefer_seq(d, m, a.B, a.nchunk, false) && seq_supported(d, m, disc_mode, nblk))
    return seq_filter(st, d, m, disc_mode, nblk, h_identity, a);
  if (m > d) return set_error(PHYSS_ERR_UNSUPPORTED, "filter: m > d is not supported");
  if (!force_grp(d) && rt_supported(d, m)) return rt_filter(st, d, m, disc_mode, nblk, h_identity, a);
  return grp_filter(st, d, m, disc_mode, nblk, h_identity, a);
}

// true when run_filter_any sends this shape to the register (physs_seq*) or register-tile (physs_rt*) kernels: the
// families that implement the pass_changed / prev_changed early-out of the fix-up passes
bool run_filter_is_seq(int d, int m, int disc_mode, int nblk, int64_t B, int64_t nchunk) {
  if (!force_grp(d) && prefer_seq(d, m, B, nchunk, false) && seq_supported(d, m, disc_mode, nblk)) return true;
  return m <= d && !force_grp(d) && rt_supported(d, m);          // the register-tile kernels implement it as well
}

int run_smooth_any(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a) {
  if (disc_mode == PHYSS_DISC_IWP && !seq_supported(d, mo == 0 ? d : mo, disc_mode, nblk))
    return set_error(PHYSS_ERR_UNSUPPORTED, "DISC_IWP: one block of state dim 2..4 (use DISC_GIVEN otherwise)");
  if (!force_grp(d) && prefer_seq(d, 1, a.B, a.nchunk, true) && seq_supported(d, mo == 0 ? d : mo, disc_mode, nblk))
    return seq_smooth(st, d, mo, disc_mode, nblk, a);
  if (!force_grp(d) && rt_supported(d, mo == 0 ? d : mo)) return rt_smooth(st, d, mo, disc_mode, nblk, a);
  return grp_smooth(st, d, mo, disc_mode, nblk, a);
}

}  // namespace physs

using namespace physs;

static inline bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }
template <typename... Ps>
static inline bool any_misaligned(Ps... ps) { return (misaligned(ps) || ...); }

// (0, 0) = default batch-major; otherwise rows must not overlap: (bs >= T, ts = 1) or (bs = 1, ts >= B)
static inline bool step_strides_ok(int64_t B, int64_t T, int64_t bs, int64_t ts) {
  if (bs == 0 && ts == 0) return true;
  if (ts == 1 && bs >= T) return true;
  if (bs == 1 && ts >= B) return true;
  return false;
}

extern "C" {

int physs_abi_version(void) { return 14; }

const char* physs_last_error(void) { return g_err; }

int physs_kf_supported(int32_t d, int32_t m, int32_t disc_mode, int32_t nblk) {
  if (seq_supported(d, m, disc_mode, nblk)) return 1;
  if (disc_mode == PHYSS_DISC_IWP) return 0;
  if (disc_mode == PHYSS_DISC_MATERN) {
    if (nblk < 1 || d % nblk != 0 || d / nblk > 4) return 0;
  }
  return grp_supported(d, m) ? 1 : 0;
}

int64_t physs_kf_wave_series(int32_t d, int32_t disc_mode, int32_t nblk) {
  if (d < 1) return 0;
  if (disc_mode == PHYSS_DISC_MATERN && (nblk < 1 || d % nblk != 0)) return 0;
  // only the register (d <= 4) and register-tile (d <= 32) families answer the query without launching; the
  // runtime-sized fallback has no fixed wave
  if (d > 32 || force_grp(d)) return 0;
  int64_t wave = 0;
  SeqSmoothArgs a{};
  a.B = 1; a.T = 1; a.sbs = 1; a.sts = 1;
  a.wave_out = &wave;
  if (run_smooth_any(nullptr, d, 0, disc_mode, nblk, a) != PHYSS_OK) return 0;
  return wave;
}

// ---- shared argument checking / packing of the filter and smoother families
#define FILTER_PARAMS                                                                             \
  void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride, int32_t d, int32_t m, \
      int32_t disc_mode, int32_t nblk, const double* A, int64_t A_bstride, const double* Q,      \
      int64_t Q_bstride, const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride, \
      const double* Pinf, int64_t Pinf_bstride, const double* m0, int64_t m0_bstride, const double* P0, \
      int64_t P0_bstride, const double* H, int64_t H_bstride, const double* Y, const double* R,   \
      int64_t R_bstride, int64_t R_tstride, double jitter
#define FILTER_ARGS                                                                               \
  stream, B, T, step_bstride, step_tstride, d, m, disc_mode, nblk, A, A_bstride, Q, Q_bstride, lam, \
      lam_bstride, dt, dt_bstride, Pinf, Pinf_bstride, m0, m0_bstride, P0, P0_bstride, H, H_bstride, Y, R, \
      R_bstride, R_tstride, jitter
#define SMOOTH_PARAMS                                                                             \
  void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride, int32_t d,     \
      int32_t disc_mode, int32_t nblk, const double* A, int64_t A_bstride, const double* Q,      \
      int64_t Q_bstride, const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride, \
      const double* Pinf, int64_t Pinf_bstride, const double* mf, const double* Pf, const double* Hout, \
      int32_t mo, double jitter
#define SMOOTH_ARGS                                                                               \
  stream, B, T, step_bstride, step_tstride, d, disc_mode, nblk, A, A_bstride, Q, Q_bstride, lam, lam_bstride, \
      dt, dt_bstride, Pinf, Pinf_bstride, mf, Pf, Hout, mo, jitter

static int pack_filter(FILTER_PARAMS, double* mf, double* Pf, double* lml, double* lml_k, SeqFilterArgs& a) {
  (void)stream;
  if (B < 0 || T < 1 || d < 1 || m < 1) return set_error(PHYSS_ERR_BAD_ARG, "filter: bad sizes");
  if (B == 0) return PHYSS_OK;
  if (!dt || !m0 || !P0 || !Y || !R || !mf || !Pf || !lml)
    return set_error(PHYSS_ERR_BAD_ARG, "filter: null required pointer");
  if (!H && m != d) return set_error(PHYSS_ERR_BAD_ARG, "filter: H == NULL (identity) needs m == d");
  if (any_misaligned(A, Q, Pinf, m0, P0, Y, R, mf, Pf))
    return set_error(PHYSS_ERR_BAD_ARG, "filter: matrix/vector pointers must be 16-byte aligned");
  if (disc_mode == PHYSS_DISC_GIVEN) {
    if (!A || !Q) return set_error(PHYSS_ERR_BAD_ARG, "filter: DISC_GIVEN needs A and Q");
  } else if (disc_mode == PHYSS_DISC_MATERN) {
    if (!lam || !Pinf || nblk < 1 || d % nblk != 0)
      return set_error(PHYSS_ERR_BAD_ARG, "filter: DISC_MATERN needs lam, Pinf and nblk | d");
  } else if (disc_mode == PHYSS_DISC_IWP) {
    if (!lam || nblk < 1 || d % nblk != 0)
      return set_error(PHYSS_ERR_BAD_ARG, "filter: DISC_IWP needs lam (spectral densities) and nblk | d");
  } else {
    return set_error(PHYSS_ERR_BAD_ARG, "filter: unknown disc_mode");
  }
  if (!step_strides_ok(B, T, step_bstride, step_tstride))
    return set_error(PHYSS_ERR_BAD_ARG, "filter: step strides must be (0,0), batch-major (>=T,1) or time-major (1,>=B)");
  a = SeqFilterArgs{};
  a.B = B; a.T = T;
  a.sbs = step_bstride ? step_bstride : T; a.sts = step_bstride ? step_tstride : 1;
  a.A = A; a.A_bs = A_bstride; a.Q = Q; a.Q_bs = Q_bstride;
  a.lam = lam; a.lam_bs = lam_bstride; a.dt = dt; a.dt_bs = dt_bstride;
  a.Pinf = Pinf; a.Pinf_bs = Pinf_bstride; a.m0 = m0; a.m0_bs = m0_bstride;
  a.P0 = P0; a.P0_bs = P0_bstride; a.H = H; a.H_bs = H_bstride;
  a.Y = Y; a.R = R; a.R_bs = R_bstride; a.R_ts = R_tstride;
  a.jitter = jitter; a.mf = mf; a.Pf = Pf; a.lml = lml; a.lml_k = lml_k;
  return PHYSS_OK;
}

static int pack_smooth(SMOOTH_PARAMS, double* ms, double* Ps, SeqSmoothArgs& a) {
  (void)stream;
  if (B < 0 || T < 1 || d < 1 || mo < 0) return set_error(PHYSS_ERR_BAD_ARG, "smoother: bad sizes");
  if (B == 0) return PHYSS_OK;
  if (!dt || !mf || !Pf) return set_error(PHYSS_ERR_BAD_ARG, "smoother: null required pointer");
  if (mo > 0 && !Hout) return set_error(PHYSS_ERR_BAD_ARG, "smoother: mo > 0 needs Hout");
  if (any_misaligned(A, Q, Pinf, mf, Pf, ms, Ps))
    return set_error(PHYSS_ERR_BAD_ARG, "smoother: matrix/vector pointers must be 16-byte aligned");
  if (disc_mode == PHYSS_DISC_GIVEN) {
    if (!A || !Q) return set_error(PHYSS_ERR_BAD_ARG, "smoother: DISC_GIVEN needs A and Q");
  } else if (disc_mode == PHYSS_DISC_MATERN) {
    if (!lam || !Pinf || nblk < 1 || d % nblk != 0)
      return set_error(PHYSS_ERR_BAD_ARG, "smoother: DISC_MATERN needs lam, Pinf and nblk | d");
  } else if (disc_mode == PHYSS_DISC_IWP) {
    if (!lam || nblk < 1 || d % nblk != 0)
      return set_error(PHYSS_ERR_BAD_ARG, "smoother: DISC_IWP needs lam (spectral densities) and nblk | d");
  } else {
    return set_error(PHYSS_ERR_BAD_ARG, "smoother: unknown disc_mode");
  }
  if (!step_strides_ok(B, T, step_bstride, step_tstride))
    return set_error(PHYSS_ERR_BAD_ARG, "smoother: step strides must be (0,0), batch-major (>=T,1) or time-major (1,>=B)");
  a = SeqSmoothArgs{};
  a.B = B; a.T = T;
  a.sbs = step_bstride ? step_bstride : T; a.sts = step_bstride ? step_tstride : 1;
  a.A = A; a.A_bs = A_bstride; a.Q = Q; a.Q_bs = Q_bstride;
  a.lam = lam; a.lam_bs = lam_bstride; a.dt = dt; a.dt_bs = dt_bstride;
  a.Pinf = Pinf; a.Pinf_bs = Pinf_bstride; a.mf = mf; a.Pf = Pf;
  a.Hout = (mo > 0) ? Hout : nullptr; a.jitter = jitter; a.ms = ms; a.Ps = Ps;
  return PHYSS_OK;
}

static int check_chunk(int64_t T, int64_t chunk_len, const void* ws) {
  if (chunk_len < 1 || chunk_len > T) return set_error(PHYSS_ERR_BAD_ARG, "pscan: need 1 <= chunk_len <= T");
  if (!ws || misaligned(ws)) return set_error(PHYSS_ERR_BAD_ARG, "pscan: workspace missing or misaligned");
  return PHYSS_OK;
}

int physs_kf_filter_f64(FILTER_PARAMS, double* mf, double* Pf, double* lml, double* lml_k) {
  SeqFilterArgs a;
  int rc = pack_filter(FILTER_ARGS, mf, Pf, lml, lml_k, a);
  if (rc || B == 0) return rc;
  return run_filter_any((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a);
}

int physs_kf_filter_smooth_f64(FILTER_PARAMS, const double* A_smooth, const double* Q_smooth, const double* dt_smooth,
                               int64_t dt_smooth_bstride, const double* Hout, int32_t mo, double* mf, double* Pf,
                               double* lml, double* lml_k, double* ms, double* Ps) {
  int rc = physs_kf_filter_f64(FILTER_ARGS, mf, Pf, lml, lml_k);
  if (rc) return rc;
  // DISC_GIVEN: the smoother's transitions follow ITS dt convention (A_smooth[k] = expm(F dt_smooth[k]))
  return physs_rts_smooth_f64(stream, B, T, step_bstride, step_tstride, d, disc_mode, nblk,
                              disc_mode == PHYSS_DISC_GIVEN ? A_smooth : A, A_bstride,
                              disc_mode == PHYSS_DISC_GIVEN ? Q_smooth : Q, Q_bstride, lam, lam_bstride, dt_smooth,
                              dt_smooth_bstride, Pinf, Pinf_bstride, mf, Pf, Hout, mo, jitter, ms, Ps);
}

// rows of the packed hand-over buffer: [m | upper triangle of P], padded to an even count (PackedRow<D>::N)
static inline int64_t packed_row_doubles(int32_t d) { return (d + d * (d + 1) / 2 + 1) & ~1; }

int physs_kf_filter_smooth_packed_supported(int32_t d, int32_t m, int32_t disc_mode, int32_t nblk) {
  return (d >= 1 && d <= 4 && seq_supported(d, m, disc_mode, nblk)) ? 1 : 0;
}

int64_t physs_kf_filter_smooth_packed_ws_bytes(int64_t B, int64_t T, int64_t step_tstride, int32_t d) {
  if (B < 0 || T < 1 || d < 1) return 0;
  const int64_t ts = step_tstride > B ? step_tstride : B;
  return ts * T * packed_row_doubles(d) * (int64_t)sizeof(double);
}

int physs_kf_filter_smooth_packed_f64(FILTER_PARAMS, const double* A_smooth, const double* Q_smooth,
                                      const double* dt_smooth, int64_t dt_smooth_bstride, const double* Hout,
                                      int32_t mo, void* ws, int64_t ws_bytes, double* lml, double* lml_k, double* ms,
                                      double* Ps) {
  if (!physs_kf_filter_smooth_packed_supported(d, m, disc_mode, nblk))
    return set_error(PHYSS_ERR_UNSUPPORTED, "packed filter + smoother: register kernels, d <= 4");
  if (step_bstride != 1 || step_tstride < B)
    return set_error(PHYSS_ERR_UNSUPPORTED, "packed filter + smoother: time-major steps (strides (1, >= B))");
  if (!ws || misaligned(ws) || ws_bytes < physs_kf_filter_smooth_packed_ws_bytes(B, T, step_tstride, d))
    return set_error(PHYSS_ERR_BAD_ARG, "packed filter + smoother: workspace missing, misaligned or too small");
  if (!dt_smooth || !ms || !Ps) return set_error(PHYSS_ERR_BAD_ARG, "packed filter + smoother: null required pointer");
  double* pk = static_cast<double*>(ws);
  SeqFilterArgs f;
  // pack_filter insists on (mf, Pf): the packed rows stand in for both (never written through these pointers)
  int rc = pack_filter(FILTER_ARGS, pk, pk, lml, lml_k, f);
  if (rc || B == 0) return rc;
  f.mf = nullptr; f.Pf = nullptr; f.pk = pk;
  rc = seq_filter((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, f);
  if (rc) return rc;
  SeqSmoothArgs s;
  rc = pack_smooth(stream, B, T, step_bstride, step_tstride, d, disc_mode, nblk,
                   disc_mode == PHYSS_DISC_GIVEN ? A_smooth : A, A_bstride,
                   disc_mode == PHYSS_DISC_GIVEN ? Q_smooth : Q, Q_bstride, lam, lam_bstride, dt_smooth,
                   dt_smooth_bstride, Pinf, Pinf_bstride, pk, pk, Hout, mo, jitter, ms, Ps, s);
  if (rc) return rc;
  s.mf = nullptr; s.Pf = nullptr; s.pk = pk;
  return seq_smooth((cudaStream_t)stream, d, Hout ? mo : 0, disc_mode, nblk, s);
}

int physs_kf_filter_colloc_f64(FILTER_PARAMS, int32_t pc, const double* res_w, int32_t n_terms,
                               const int32_t* term_out, const int32_t* term_kind, const int32_t* term_idx,
                               const double* term_coef, const double* forcing, const double* y_pseudo,
                               const double* boundary, int32_t observe_data, double* mf, double* Pf, double* lml,
                               double* lml_k) {
  SeqFilterArgs a;
  int rc = pack_filter(FILTER_ARGS, mf, Pf, lml, lml_k, a);
  if (rc || B == 0) return rc;
  if (n_terms > 0 && (!term_out || !term_kind || !term_idx || !term_coef))
    return set_error(PHYSS_ERR_BAD_ARG, "collocation filter: null residual term table");
  if (misaligned(boundary)) return set_error(PHYSS_ERR_BAD_ARG, "collocation filter: boundary must be 16-byte aligned");
  return colloc_filter((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a, pc, res_w, n_terms, term_out,
                       term_kind, term_idx, term_coef, forcing, y_pseudo, boundary, observe_data);
}

int physs_spd_inverse_f64(void* stream, int64_t N, int32_t D, const double* A, double jitter, double* out) {
  if (N < 0 || D < 1) return set_error(PHYSS_ERR_BAD_ARG, "spd inverse: bad sizes");
  if (N > 0 && (!A || !out)) return set_error(PHYSS_ERR_BAD_ARG, "spd inverse: null pointer");
  return spd_inverse((cudaStream_t)stream, N, D, A, jitter, out);
}

int physs_kf_vjp_supported(int32_t d, int32_t m, int32_t disc_mode, int32_t nblk) {
  return (vjp_supported(d, m, disc_mode, nblk) || grp_vjp_supported(d, m, disc_mode)) ? 1 : 0;
}

int physs_kf_filter_vjp_f64(FILTER_PARAMS, const double* mf, const double* Pf, const double* g_lml, double* gA,
                            double* gQ, double* glam, double* gPinf, double* gH, double* gR_step, double* gR_sum,
                            double* gm0, double* gP0) {
  SeqFilterArgs a;
  double dummy = 0.0;                       // pack_filter insists on an lml pointer; the reverse pass writes none
  int rc = pack_filter(FILTER_ARGS, const_cast<double*>(mf), const_cast<double*>(Pf), &dummy, nullptr, a);
  if (rc || B == 0) return rc;
  a.lml = nullptr;
  if (disc_mode == PHYSS_DISC_GIVEN ? (!gA || !gQ) : (!glam || !gPinf))
    return set_error(PHYSS_ERR_BAD_ARG, "filter vjp: missing gradient output for this disc_mode");
  VjpOut o{};
  o.g_lml = g_lml; o.gA = gA; o.gQ = gQ; o.glam = glam; o.gPinf = gPinf; o.gH = gH;
  o.gR_step = gR_step; o.gR_sum = gR_sum; o.gm0 = gm0; o.gP0 = gP0;
  if (!vjp_supported(d, m, disc_mode, nblk) && grp_vjp_supported(d, m, disc_mode))
    return grp_kf_vjp((cudaStream_t)stream, d, m, H == nullptr, a, o);       // general (d, m): lane groups
  return kf_vjp((cudaStream_t)stream, d, m, disc_mode, nblk, a, o);
}

int physs_rts_smooth_f64(SMOOTH_PARAMS, double* ms, double* Ps) {
  SeqSmoothArgs a;
  int rc = pack_smooth(SMOOTH_ARGS, ms, Ps, a);
  if (rc || B == 0) return rc;
  if (!ms || !Ps) return set_error(PHYSS_ERR_BAD_ARG, "smoother: null output pointer");
  return run_smooth_any((cudaStream_t)stream, d, Hout ? mo : 0, disc_mode, nblk, a);
}

int physs_sum_steps_f64(void* stream, int64_t B, int64_t T, int64_t bstride, int64_t tstride, const double* x,
                        const double* sub, double* scratch, double* out) {
  if (B < 0 || T < 0) return set_error(PHYSS_ERR_BAD_ARG, "sum over steps: bad sizes");
  if (B > 0 && T > 0 && (!x || !scratch || !out)) return set_error(PHYSS_ERR_BAD_ARG, "sum over steps: null pointer");
  return sum_steps((cudaStream_t)stream, B, T, bstride, tstride, x, sub, scratch, out);
}

int64_t physs_pscan_workspace_bytes(int64_t B, int64_t T, int32_t d, int64_t chunk_len) {
  if (B < 1 || T < 1 || d < 1 || chunk_len < 1) return 0;
  return 8 * pscan_workspace_doubles(B, T, d, chunk_len);
}

int physs_pscan_filter_f64(FILTER_PARAMS, int64_t chunk_len, int32_t polish, double delta, int32_t patience,
                           void* ws, double* mf, double* Pf, double* lml, double* lml_k, int32_t* status) {
  SeqFilterArgs a;
  int rc = pack_filter(FILTER_ARGS, mf, Pf, lml, lml_k, a);
  if (rc || B == 0) return rc;
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  rc = pscan_filter_local((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a, chunk_len, (double*)ws, nullptr);
  if (rc) return rc;
  return pscan_filter_finish((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a, chunk_len, (double*)ws,
                             false, nullptr, nullptr, polish, delta, patience, status);
}

int physs_pscan_filter_spec_f64(FILTER_PARAMS, int64_t chunk_len, int64_t warm, int32_t polish, double delta,
                                int32_t patience, void* ws, double* mf, double* Pf, double* lml, double* lml_k,
                                int32_t* status) {
  SeqFilterArgs a;
  int rc = pack_filter(FILTER_ARGS, mf, Pf, lml, lml_k, a);
  if (rc || B == 0) return rc;
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  return pscan_filter_spec((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a, chunk_len, warm, polish,
                           delta, patience, (double*)ws, status);
}

int physs_pscan_smooth_spec_f64(SMOOTH_PARAMS, int64_t chunk_len, int64_t warm, int32_t polish, double delta,
                                int32_t patience, void* ws, double* ms, double* Ps, int32_t* status) {
  SeqSmoothArgs a;
  int rc = pack_smooth(SMOOTH_ARGS, ms, Ps, a);
  if (rc || B == 0) return rc;
  if (!ms || !Ps) return set_error(PHYSS_ERR_BAD_ARG, "smoother: null output pointer");
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  return pscan_smooth_spec((cudaStream_t)stream, d, Hout ? mo : 0, disc_mode, nblk, a, chunk_len, warm, polish,
                           delta, patience, (double*)ws, status);
}

int physs_pscan_filter_local_f64(FILTER_PARAMS, int64_t chunk_len, void* ws, double* total) {
  SeqFilterArgs a;
  double dummy;   // outputs are not touched by the local phase
  int rc = pack_filter(FILTER_ARGS, (double*)ws, (double*)ws, &dummy, nullptr, a);
  if (rc || B == 0) return rc;
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  if (!total) return set_error(PHYSS_ERR_BAD_ARG, "pscan local: total is required");
  return pscan_filter_local((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a, chunk_len, (double*)ws, total);
}

int physs_pscan_filter_finish_f64(FILTER_PARAMS, int64_t chunk_len, int32_t polish, double delta, int32_t patience,
                                  void* ws, const double* start_m, const double* start_P, double* mf, double* Pf,
                                  double* lml, double* lml_k, int32_t* status) {
  SeqFilterArgs a;
  int rc = pack_filter(FILTER_ARGS, mf, Pf, lml, lml_k, a);
  if (rc || B == 0) return rc;
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  return pscan_filter_finish((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a, chunk_len, (double*)ws,
                             true, start_m, start_P, polish, delta, patience, status);
}

int physs_pscan_filter_fold_f64(void* stream, int64_t B, int32_t d, int64_t K, const double* totals,
                                const double* m0, int64_t m0_bstride, const double* P0, int64_t P0_bstride,
                                double* m_out, double* P_out) {
  if (B < 0 || d < 1 || K < 0) return set_error(PHYSS_ERR_BAD_ARG, "pscan fold: bad sizes");
  if (B == 0) return PHYSS_OK;
  if (!m0 || !P0 || !m_out || !P_out || (K > 0 && !totals))
    return set_error(PHYSS_ERR_BAD_ARG, "pscan fold: null required pointer");
  return pscan_filter_fold((cudaStream_t)stream, d, B, K, totals, m0, m0_bstride, P0, P0_bstride, m_out, P_out);
}

int physs_pscan_smooth_f64(SMOOTH_PARAMS, int64_t chunk_len, void* ws, double* ms, double* Ps) {
  SeqSmoothArgs a;
  int rc = pack_smooth(SMOOTH_ARGS, ms, Ps, a);
  if (rc || B == 0) return rc;
  if (!ms || !Ps) return set_error(PHYSS_ERR_BAD_ARG, "smoother: null output pointer");
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  rc = pscan_smooth_local((cudaStream_t)stream, d, disc_mode, nblk, a, chunk_len, (double*)ws, nullptr);
  if (rc) return rc;
  return pscan_smooth_finish((cudaStream_t)stream, d, Hout ? mo : 0, disc_mode, nblk, a, chunk_len, (double*)ws,
                             nullptr, nullptr);
}

int physs_pscan_smooth_local_f64(SMOOTH_PARAMS, int64_t chunk_len, void* ws, double* total) {
  SeqSmoothArgs a;
  int rc = pack_smooth(SMOOTH_ARGS, nullptr, nullptr, a);
  if (rc || B == 0) return rc;
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  if (!total) return set_error(PHYSS_ERR_BAD_ARG, "pscan local: total is required");
  return pscan_smooth_local((cudaStream_t)stream, d, disc_mode, nblk, a, chunk_len, (double*)ws, total);
}

int physs_pscan_smooth_finish_f64(SMOOTH_PARAMS, int64_t chunk_len, void* ws, const double* start_m,
                                  const double* start_P, double* ms, double* Ps) {
  SeqSmoothArgs a;
  int rc = pack_smooth(SMOOTH_ARGS, ms, Ps, a);
  if (rc || B == 0) return rc;
  if (!ms || !Ps) return set_error(PHYSS_ERR_BAD_ARG, "smoother: null output pointer");
  if ((rc = check_chunk(T, chunk_len, ws))) return rc;
  return pscan_smooth_finish((cudaStream_t)stream, d, Hout ? mo : 0, disc_mode, nblk, a, chunk_len, (double*)ws,
                             start_m, start_P);
}

int physs_pscan_smooth_fold_f64(void* stream, int64_t B, int32_t d, int64_t K, const double* totals,
                                const double* m_end, const double* P_end, double* m_out, double* P_out) {
  if (B < 0 || d < 1 || K < 0) return set_error(PHYSS_ERR_BAD_ARG, "pscan fold: bad sizes");
  if (B == 0) return PHYSS_OK;
  if (!m_end || !P_end || !m_out || !P_out || (K > 0 && !totals))
    return set_error(PHYSS_ERR_BAD_ARG, "pscan fold: null required pointer");
  return pscan_smooth_fold((cudaStream_t)stream, d, B, K, totals, m_end, P_end, m_out, P_out);
}

static int cvi_dispatch(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik, bool update,
                        const CviArgs& a) {
  if (N < 0 || D < 1) return set_error(PHYSS_ERR_BAD_ARG, "cvi: bad sizes");
  if (N == 0) return PHYSS_OK;
  if (lik != 3 && (P < 1 || P > D)) return set_error(PHYSS_ERR_BAD_ARG, "cvi: need 1 <= P <= D");
  if (!a.W && lik != 3 && P != D) return set_error(PHYSS_ERR_BAD_ARG, "cvi: W == NULL (identity) needs P == D");
  if ((lik == 1 || lik == 2) && (a.K < 1 || !a.ghx || !a.ghw))
    return set_error(PHYSS_ERR_BAD_ARG, "cvi: Gauss-Hermite likelihoods need K, ghx, ghw");
  if (cvi_reg_supported(D, lik == 3 ? 1 : P)) return cvi_reg_run((cudaStream_t)stream, D, P, lik, update, a);
  return cvi_grp_run((cudaStream_t)stream, D, P, lik, update, a);
}

static int cvi_natgrad_step(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik, const double* Ytil,
                            const double* Vtil, const double* q_mu, const double* q_var, const double* y,
                            const double* W, const double* noise, int64_t noise_stride, double lik_param, int32_t K,
                            const double* ghx, const double* ghw, const double* dm_in, const double* dS_in, double beta,
                            double ng_jitter, double* Ytil_out, double* Vtil_out, double* ell_out, int prec);

int physs_cvi_natgrad_step_f64(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik,
                               const double* Ytil, const double* Vtil,
                               const double* q_mu, const double* q_var,
                               const double* y, const double* W,
                               const double* noise, int64_t noise_stride,
                               double lik_param, int32_t K, const double* ghx, const double* ghw,
                               const double* dm_in, const double* dS_in,
                               double beta, double ng_jitter,
                               double* Ytil_out, double* Vtil_out, double* ell_out) {
  return cvi_natgrad_step(stream, N, D, P, lik, Ytil, Vtil, q_mu, q_var, y, W, noise, noise_stride, lik_param, K, ghx,
                          ghw, dm_in, dS_in, beta, ng_jitter, Ytil_out, Vtil_out, ell_out, 0);
}

int physs_cvi_natgrad_step_prec_f64(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik,
                                    const double* Ytil, const double* Ptil,
                                    const double* q_mu, const double* q_var,
                                    const double* y, const double* W,
                                    const double* noise, int64_t noise_stride,
                                    double lik_param, int32_t K, const double* ghx, const double* ghw,
                                    const double* dm_in, const double* dS_in,
                                    double beta, double ng_jitter,
                                    double* Ytil_out, double* Ptil_out, double* ell_out) {
  return cvi_natgrad_step(stream, N, D, P, lik, Ytil, Ptil, q_mu, q_var, y, W, noise, noise_stride, lik_param, K, ghx,
                          ghw, dm_in, dS_in, beta, ng_jitter, Ytil_out, Ptil_out, ell_out, 1);
}

static int cvi_natgrad_step(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik, const double* Ytil,
                            const double* Vtil, const double* q_mu, const double* q_var, const double* y,
                            const double* W, const double* noise, int64_t noise_stride, double lik_param, int32_t K,
                            const double* ghx, const double* ghw, const double* dm_in, const double* dS_in, double beta,
                            double ng_jitter, double* Ytil_out, double* Vtil_out, double* ell_out, int prec) {
  if (N > 0 && (!Ytil || !Vtil || !q_mu || !q_var || !Ytil_out || !Vtil_out))
    return set_error(PHYSS_ERR_BAD_ARG, "cvi step: null required pointer");
  if (lik == 3 && N > 0 && (!dm_in || !dS_in)) return set_error(PHYSS_ERR_BAD_ARG, "cvi step: LIK_GIVEN needs dm, dS");
  if (lik != 3 && N > 0 && !y) return set_error(PHYSS_ERR_BAD_ARG, "cvi step: needs data y");
  if (lik == 0 && N > 0 && !noise) return set_error(PHYSS_ERR_BAD_ARG, "cvi step: Gaussian likelihood needs noise");
  CviArgs a{};
  a.N = N; a.Yt = Ytil; a.Vt = Vtil; a.qm = q_mu; a.qS = q_var; a.y = y; a.W = W;
  a.noise = noise; a.noise_stride = noise_stride; a.lik_param = lik_param; a.K = K; a.ghx = ghx; a.ghw = ghw;
  a.dm_in = dm_in; a.dS_in = dS_in; a.beta = beta; a.ngj = ng_jitter; a.prec = prec;
  a.Yn = Ytil_out; a.Vn = Vtil_out; a.ell = ell_out; a.dm_out = nullptr; a.dS_out = nullptr;
  return cvi_dispatch(stream, N, D, P, lik, true, a);
}

int physs_cvi_ell_f64(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik,
                      const double* q_mu, const double* q_var,
                      const double* y, const double* W,
                      const double* noise, int64_t noise_stride,
                      double lik_param, int32_t K, const double* ghx, const double* ghw,
                      double* ell_out, double* dm_out, double* dS_out) {
  if (N > 0 && (!q_mu || !q_var || !y)) return set_error(PHYSS_ERR_BAD_ARG, "cvi ell: null required pointer");
  if (lik < 0 || lik > 2) return set_error(PHYSS_ERR_BAD_ARG, "cvi ell: likelihood must be 0, 1 or 2");
  if (lik == 0 && N > 0 && !noise) return set_error(PHYSS_ERR_BAD_ARG, "cvi ell: Gaussian likelihood needs noise");
  CviArgs a{};
  a.N = N; a.qm = q_mu; a.qS = q_var; a.y = y; a.W = W;
  a.noise = noise; a.noise_stride = noise_stride; a.lik_param = lik_param; a.K = K; a.ghx = ghx; a.ghw = ghw;
  a.ell = ell_out; a.dm_out = dm_out; a.dS_out = dS_out;
  return cvi_dispatch(stream, N, D, P, lik, false, a);
}

}  // extern "C"
