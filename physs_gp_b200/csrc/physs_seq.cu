// physs_seq.cu -- size dispatch for the register-resident sequential filter / smoother
// (one thread per series, state dim d <= 4).  The kernels live in physs_seq_impl.cuh and are
// instantiated one (d, block size, discretisation) triple per translation unit (physs_seq_d*.cu) so
// that the library builds in parallel.
#include "physs_internal.h"

namespace physs {

int seq_filter_d1s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d1s1m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d1s1g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d1s1g(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d2s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d2s1m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d2s2m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d2s2m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d2s2g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d2s2g(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d3s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d3s1m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d3s3m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d3s3m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d3s3g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d3s3g(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d4s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d4s1m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d4s2m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d4s2m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d4s4m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d4s4m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d4s4g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_filter_d8s4m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d8s4m(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_smooth_d4s4g(cudaStream_t st, const SeqSmoothArgs& a, int mo);

int seq_filter_d2w(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d2w(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d3w(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d3w(cudaStream_t st, const SeqSmoothArgs& a, int mo);
int seq_filter_d4w(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid);
int seq_smooth_d4w(cudaStream_t st, const SeqSmoothArgs& a, int mo);

int seq_filter_summary_d1s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d1s1m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d1s1g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d1s1g(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d2s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d2s1m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d2s2m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d2s2m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d2s2g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d2s2g(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d3s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d3s1m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d3s3m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d3s3m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d3s3g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d3s3g(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d4s1m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d4s1m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d4s2m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d4s2m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d4s4m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d4s4m(cudaStream_t st, const SeqSmoothArgs& a, double* elems);
int seq_filter_summary_d4s4g(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems);
int seq_smooth_summary_d4s4g(cudaStream_t st, const SeqSmoothArgs& a, double* elems);

bool seq_supported(int d, int m, int disc_mode, int nblk) {
  // d = 8 as two Matern-7/2 blocks (m <= 4 or full-state m = 8): thread-per-series with local-memory tiles
  if (d == 8 && disc_mode == PHYSS_DISC_MATERN && nblk == 2 && (m == 8 || (m >= 1 && m <= 4))) return true;
  if (d < 1 || d > 4 || m < 1 || m > d) return false;
  if (disc_mode == PHYSS_DISC_GIVEN) return true;
  if (disc_mode == PHYSS_DISC_IWP) return nblk == 1 && d >= 2;      // one IWP(q) block, q = d - 1 in 1..3
  if (disc_mode != PHYSS_DISC_MATERN || nblk <= 0 || d % nblk != 0) return false;
  const int s = d / nblk;
  return s == 1 || s == d || (d == 4 && s == 2);
}

int seq_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
               const SeqFilterArgs& a) {
  if (!seq_supported(d, m, disc_mode, nblk))
    return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter: no specialisation for this (d, m, blocks)");
  if (disc_mode == PHYSS_DISC_IWP) {
    if (a.nchunk > 0) return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter: DISC_IWP has no chunk mode");
    if (d == 2) return seq_filter_d2w(st, a, m, h_identity);
    if (d == 3) return seq_filter_d3w(st, a, m, h_identity);
    return seq_filter_d4w(st, a, m, h_identity);
  }
  const int s = (disc_mode == PHYSS_DISC_GIVEN) ? d : d / nblk;
  const bool g = (disc_mode == PHYSS_DISC_GIVEN);
  if (d == 1 && s == 1 && g == false) return seq_filter_d1s1m(st, a, m, h_identity);
  if (d == 1 && s == 1 && g == true) return seq_filter_d1s1g(st, a, m, h_identity);
  if (d == 2 && s == 1 && g == false) return seq_filter_d2s1m(st, a, m, h_identity);
  if (d == 2 && s == 2 && g == false) return seq_filter_d2s2m(st, a, m, h_identity);
  if (d == 2 && s == 2 && g == true) return seq_filter_d2s2g(st, a, m, h_identity);
  if (d == 3 && s == 1 && g == false) return seq_filter_d3s1m(st, a, m, h_identity);
  if (d == 3 && s == 3 && g == false) return seq_filter_d3s3m(st, a, m, h_identity);
  if (d == 3 && s == 3 && g == true) return seq_filter_d3s3g(st, a, m, h_identity);
  if (d == 4 && s == 1 && g == false) return seq_filter_d4s1m(st, a, m, h_identity);
  if (d == 4 && s == 2 && g == false) return seq_filter_d4s2m(st, a, m, h_identity);
  if (d == 4 && s == 4 && g == false) return seq_filter_d4s4m(st, a, m, h_identity);
  if (d == 4 && s == 4 && g == true) return seq_filter_d4s4g(st, a, m, h_identity);
  if (d == 8 && s == 4 && g == false) return seq_filter_d8s4m(st, a, m, h_identity);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter: unreachable");
}

int seq_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a) {
  if (!seq_supported(d, mo == 0 ? d : mo, disc_mode, nblk))
    return set_error(PHYSS_ERR_UNSUPPORTED, "seq smoother: no specialisation for this (d, mo, blocks)");
  if (disc_mode == PHYSS_DISC_IWP) {
    if (a.nchunk > 0) return set_error(PHYSS_ERR_UNSUPPORTED, "seq smoother: DISC_IWP has no chunk mode");
    if (d == 2) return seq_smooth_d2w(st, a, mo);
    if (d == 3) return seq_smooth_d3w(st, a, mo);
    return seq_smooth_d4w(st, a, mo);
  }
  const int s = (disc_mode == PHYSS_DISC_GIVEN) ? d : d / nblk;
  const bool g = (disc_mode == PHYSS_DISC_GIVEN);
  if (d == 1 && s == 1 && g == false) return seq_smooth_d1s1m(st, a, mo);
  if (d == 1 && s == 1 && g == true) return seq_smooth_d1s1g(st, a, mo);
  if (d == 2 && s == 1 && g == false) return seq_smooth_d2s1m(st, a, mo);
  if (d == 2 && s == 2 && g == false) return seq_smooth_d2s2m(st, a, mo);
  if (d == 2 && s == 2 && g == true) return seq_smooth_d2s2g(st, a, mo);
  if (d == 3 && s == 1 && g == false) return seq_smooth_d3s1m(st, a, mo);
  if (d == 3 && s == 3 && g == false) return seq_smooth_d3s3m(st, a, mo);
  if (d == 3 && s == 3 && g == true) return seq_smooth_d3s3g(st, a, mo);
  if (d == 4 && s == 1 && g == false) return seq_smooth_d4s1m(st, a, mo);
  if (d == 4 && s == 2 && g == false) return seq_smooth_d4s2m(st, a, mo);
  if (d == 4 && s == 4 && g == false) return seq_smooth_d4s4m(st, a, mo);
  if (d == 4 && s == 4 && g == true) return seq_smooth_d4s4g(st, a, mo);
  if (d == 8 && s == 4 && g == false) return seq_smooth_d8s4m(st, a, mo);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq smoother: unreachable");
}


int seq_filter_summary(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
                       const SeqFilterArgs& a, double* elems) {
  if (!seq_supported(d, m, disc_mode, nblk) || disc_mode == PHYSS_DISC_IWP)
    return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter summary: no specialisation for this (d, m, blocks)");
  const int s = (disc_mode == PHYSS_DISC_GIVEN) ? d : d / nblk;
  const bool g = (disc_mode == PHYSS_DISC_GIVEN);
  if (d == 1 && s == 1 && g == false) return seq_filter_summary_d1s1m(st, a, m, h_identity, elems);
  if (d == 1 && s == 1 && g == true) return seq_filter_summary_d1s1g(st, a, m, h_identity, elems);
  if (d == 2 && s == 1 && g == false) return seq_filter_summary_d2s1m(st, a, m, h_identity, elems);
  if (d == 2 && s == 2 && g == false) return seq_filter_summary_d2s2m(st, a, m, h_identity, elems);
  if (d == 2 && s == 2 && g == true) return seq_filter_summary_d2s2g(st, a, m, h_identity, elems);
  if (d == 3 && s == 1 && g == false) return seq_filter_summary_d3s1m(st, a, m, h_identity, elems);
  if (d == 3 && s == 3 && g == false) return seq_filter_summary_d3s3m(st, a, m, h_identity, elems);
  if (d == 3 && s == 3 && g == true) return seq_filter_summary_d3s3g(st, a, m, h_identity, elems);
  if (d == 4 && s == 1 && g == false) return seq_filter_summary_d4s1m(st, a, m, h_identity, elems);
  if (d == 4 && s == 2 && g == false) return seq_filter_summary_d4s2m(st, a, m, h_identity, elems);
  if (d == 4 && s == 4 && g == false) return seq_filter_summary_d4s4m(st, a, m, h_identity, elems);
  if (d == 4 && s == 4 && g == true) return seq_filter_summary_d4s4g(st, a, m, h_identity, elems);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter summary: unreachable");
}

int seq_smooth_summary(cudaStream_t st, int d, int disc_mode, int nblk, const SeqSmoothArgs& a, double* elems) {
  if (!seq_supported(d, d, disc_mode, nblk) || disc_mode == PHYSS_DISC_IWP)
    return set_error(PHYSS_ERR_UNSUPPORTED, "seq smoother summary: no specialisation for this (d, blocks)");
  const int s = (disc_mode == PHYSS_DISC_GIVEN) ? d : d / nblk;
  const bool g = (disc_mode == PHYSS_DISC_GIVEN);
  if (d == 1 && s == 1 && g == false) return seq_smooth_summary_d1s1m(st, a, elems);
  if (d == 1 && s == 1 && g == true) return seq_smooth_summary_d1s1g(st, a, elems);
  if (d == 2 && s == 1 && g == false) return seq_smooth_summary_d2s1m(st, a, elems);
  if (d == 2 && s == 2 && g == false) return seq_smooth_summary_d2s2m(st, a, elems);
  if (d == 2 && s == 2 && g == true) return seq_smooth_summary_d2s2g(st, a, elems);
  if (d == 3 && s == 1 && g == false) return seq_smooth_summary_d3s1m(st, a, elems);
  if (d == 3 && s == 3 && g == false) return seq_smooth_summary_d3s3m(st, a, elems);
  if (d == 3 && s == 3 && g == true) return seq_smooth_summary_d3s3g(st, a, elems);
  if (d == 4 && s == 1 && g == false) return seq_smooth_summary_d4s1m(st, a, elems);
  if (d == 4 && s == 2 && g == false) return seq_smooth_summary_d4s2m(st, a, elems);
  if (d == 4 && s == 4 && g == false) return seq_smooth_summary_d4s4m(st, a, elems);
  if (d == 4 && s == 4 && g == true) return seq_smooth_summary_d4s4g(st, a, elems);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq smoother summary: unreachable");
}

}  // namespace physs
