// physs_rt.cu -- sequential Kalman filter / RTS smoother for state dims 5 .. 32: one lane group per series,
// matrices resident in shared memory (zero-padded to DM = 8 / 16 / 32), products accumulated in register
// row tiles (physs_rt.cuh).  Same reference semantics (kalman_filter.py:144-241,439-485;
// rts_smoother.py:48-106,162-192), ABI and chunk / fix-up modes as physs_grp.cu, which remains the
// fallback for d > 32.
//
//   DM = 8 / 16 / 32 : G = DM lanes per series, one matrix row per lane (4 / 2 / 1 series per warp)
#include "physs_internal.h"

namespace physs {

// per-DM entry points, defined in physs_rt_d*.cu / physs_rt_sum_d*.cu
template <int DM>
int rt_filter_dm(cudaStream_t st, bool given, const SeqFilterArgs& a, int d, int m, int nblk, bool hid);
template <int DM>
int rt_smooth_dm(cudaStream_t st, bool given, const SeqSmoothArgs& a, int d, int mo, int nblk);
template <int DM>
int rt_filter_summary_dm(cudaStream_t st, bool given, const SeqFilterArgs& a, int d, int m, int nblk, bool hid,
                         int64_t cfirst, int64_t ccount, double* elems);
template <int DM>
int rt_smooth_summary_dm(cudaStream_t st, bool given, const SeqSmoothArgs& a, int d, int nblk, double* elems);
#define DECL(DM_)                                                                                          \
  template <> int rt_filter_dm<DM_>(cudaStream_t, bool, const SeqFilterArgs&, int, int, int, bool);        \
  template <> int rt_smooth_dm<DM_>(cudaStream_t, bool, const SeqSmoothArgs&, int, int, int);              \
  template <> int rt_filter_summary_dm<DM_>(cudaStream_t, bool, const SeqFilterArgs&, int, int, int, bool, \
                                            int64_t, int64_t, double*);                                    \
  template <> int rt_smooth_summary_dm<DM_>(cudaStream_t, bool, const SeqSmoothArgs&, int, int, double*);
DECL(8) DECL(16) DECL(32)
#undef DECL

// ---------------------------------------------------------------------------------------- dispatch
bool rt_supported(int d, int m) { return d >= 1 && d <= 32 && m >= 1 && m <= d; }

static int rt_check(int d, int disc_mode, int nblk) {
  if (disc_mode == PHYSS_DISC_GIVEN) return PHYSS_OK;
  const int s = (nblk > 0) ? d / nblk : 0;
  if (s < 1 || s > 4 || s * nblk != d)
    return set_error(PHYSS_ERR_UNSUPPORTED, "DISC_MATERN needs equal blocks of size 1..4");
  return PHYSS_OK;
}

// One row per lane (G = DM): the step is a chain of short dependent phases (Cholesky columns, triangular
// solves), so the per-series critical path, not the instruction count, sets the speed (measured 1.4-1.7x
// faster than two rows per lane at d = 8 / 16 / 32).
int rt_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, const SeqFilterArgs& a) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  if (d <= 8) return rt_filter_dm<8>(st, given, a, d, m, nblk, hid);
  if (d <= 16) return rt_filter_dm<16>(st, given, a, d, m, nblk, hid);
  return rt_filter_dm<32>(st, given, a, d, m, nblk, hid);
}

int rt_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  if (d <= 8) return rt_smooth_dm<8>(st, given, a, d, mo, nblk);
  if (d <= 16) return rt_smooth_dm<16>(st, given, a, d, mo, nblk);
  return rt_smooth_dm<32>(st, given, a, d, mo, nblk);
}

int rt_filter_summary(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, const SeqFilterArgs& a,
                      int64_t cfirst, int64_t ccount, double* elems) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  if (d <= 8) return rt_filter_summary_dm<8>(st, given, a, d, m, nblk, hid, cfirst, ccount, elems);
  if (d <= 16) return rt_filter_summary_dm<16>(st, given, a, d, m, nblk, hid, cfirst, ccount, elems);
  return rt_filter_summary_dm<32>(st, given, a, d, m, nblk, hid, cfirst, ccount, elems);
}

int rt_smooth_summary(cudaStream_t st, int d, int disc_mode, int nblk, const SeqSmoothArgs& a, double* elems) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  if (d <= 8) return rt_smooth_summary_dm<8>(st, given, a, d, nblk, elems);
  if (d <= 16) return rt_smooth_summary_dm<16>(st, given, a, d, nblk, elems);
  return rt_smooth_summary_dm<32>(st, given, a, d, nblk, elems);
}

}  // namespace physs
