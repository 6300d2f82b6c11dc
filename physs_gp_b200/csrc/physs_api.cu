// physs_api.cu -- the extern "C" surface of libphyss_b200.so (see include/physs_b200.h).
#include <stdio.h>
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

using namespace physs;

static inline bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }
template <typename... Ps>
static inline bool any_misaligned(Ps... ps) { return (misaligned(ps) || ...); }

extern "C" {

int physs_abi_version(void) { return 1; }

const char* physs_last_error(void) { return g_err; }

int physs_kf_supported(int32_t d, int32_t m, int32_t disc_mode, int32_t nblk) {
  if (seq_supported(d, m, disc_mode, nblk)) return 1;
  if (disc_mode == PHYSS_DISC_MATERN) {
    if (nblk < 1 || d % nblk != 0 || d / nblk > 4) return 0;
  }
  return grp_supported(d, m) ? 1 : 0;
}

int physs_kf_filter_f64(void* stream, int64_t B, int64_t T, int32_t d, int32_t m,
                        int32_t disc_mode, int32_t nblk,
                        const double* A, int64_t A_bstride,
                        const double* Q, int64_t Q_bstride,
                        const double* lam, int64_t lam_bstride,
                        const double* dt, int64_t dt_bstride,
                        const double* Pinf, int64_t Pinf_bstride,
                        const double* m0, int64_t m0_bstride,
                        const double* P0, int64_t P0_bstride,
                        const double* H, int64_t H_bstride,
                        const double* Y,
                        const double* R, int64_t R_bstride, int64_t R_tstride,
                        double jitter,
                        double* mf, double* Pf, double* lml, double* lml_k) {
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
  } else {
    return set_error(PHYSS_ERR_BAD_ARG, "filter: unknown disc_mode");
  }
  SeqFilterArgs a;
  a.B = B; a.T = T;
  a.A = A; a.A_bs = A_bstride; a.Q = Q; a.Q_bs = Q_bstride;
  a.lam = lam; a.lam_bs = lam_bstride; a.dt = dt; a.dt_bs = dt_bstride;
  a.Pinf = Pinf; a.Pinf_bs = Pinf_bstride; a.m0 = m0; a.m0_bs = m0_bstride;
  a.P0 = P0; a.P0_bs = P0_bstride; a.H = H; a.H_bs = H_bstride;
  a.Y = Y; a.R = R; a.R_bs = R_bstride; a.R_ts = R_tstride;
  a.jitter = jitter; a.mf = mf; a.Pf = Pf; a.lml = lml; a.lml_k = lml_k;
  if (seq_supported(d, m, disc_mode, nblk))
    return seq_filter((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a);
  if (m > d) return set_error(PHYSS_ERR_UNSUPPORTED, "filter: m > d is not supported");
  return grp_filter((cudaStream_t)stream, d, m, disc_mode, nblk, H == nullptr, a);
}

int physs_rts_smooth_f64(void* stream, int64_t B, int64_t T, int32_t d,
                         int32_t disc_mode, int32_t nblk,
                         const double* A, int64_t A_bstride,
                         const double* Q, int64_t Q_bstride,
                         const double* lam, int64_t lam_bstride,
                         const double* dt, int64_t dt_bstride,
                         const double* Pinf, int64_t Pinf_bstride,
                         const double* mf, const double* Pf,
                         const double* Hout, int32_t mo,
                         double jitter,
                         double* ms, double* Ps) {
  if (B < 0 || T < 1 || d < 1 || mo < 0) return set_error(PHYSS_ERR_BAD_ARG, "smoother: bad sizes");
  if (B == 0) return PHYSS_OK;
  if (!dt || !mf || !Pf || !ms || !Ps)
    return set_error(PHYSS_ERR_BAD_ARG, "smoother: null required pointer");
  if (mo > 0 && !Hout) return set_error(PHYSS_ERR_BAD_ARG, "smoother: mo > 0 needs Hout");
  if (any_misaligned(A, Q, Pinf, mf, Pf, ms, Ps))
    return set_error(PHYSS_ERR_BAD_ARG, "smoother: matrix/vector pointers must be 16-byte aligned");
  if (disc_mode == PHYSS_DISC_GIVEN) {
    if (!A || !Q) return set_error(PHYSS_ERR_BAD_ARG, "smoother: DISC_GIVEN needs A and Q");
  } else if (disc_mode == PHYSS_DISC_MATERN) {
    if (!lam || !Pinf || nblk < 1 || d % nblk != 0)
      return set_error(PHYSS_ERR_BAD_ARG, "smoother: DISC_MATERN needs lam, Pinf and nblk | d");
  } else {
    return set_error(PHYSS_ERR_BAD_ARG, "smoother: unknown disc_mode");
  }
  SeqSmoothArgs a;
  a.B = B; a.T = T;
  a.A = A; a.A_bs = A_bstride; a.Q = Q; a.Q_bs = Q_bstride;
  a.lam = lam; a.lam_bs = lam_bstride; a.dt = dt; a.dt_bs = dt_bstride;
  a.Pinf = Pinf; a.Pinf_bs = Pinf_bstride; a.mf = mf; a.Pf = Pf;
  a.Hout = (mo > 0) ? Hout : nullptr; a.jitter = jitter; a.ms = ms; a.Ps = Ps;
  const int mo_eff = Hout ? mo : 0;
  if (seq_supported(d, mo_eff == 0 ? d : mo_eff, disc_mode, nblk))
    return seq_smooth((cudaStream_t)stream, d, mo_eff, disc_mode, nblk, a);
  return grp_smooth((cudaStream_t)stream, d, mo_eff, disc_mode, nblk, a);
}

}  // extern "C"
