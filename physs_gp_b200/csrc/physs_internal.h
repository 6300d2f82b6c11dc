// physs_internal.h -- shared between the translation units of libphyss_b200.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/physs_b200.h"

namespace physs {

// thread-local last-error string behind physs_last_error()
int set_error(int code, const char* msg);
int cuda_status(cudaError_t e, const char* what);

struct SeqFilterArgs {
  int64_t B, T;
  int64_t sbs, sts;   // step strides: row (b, k) of a per-step array with n doubles is at (b*sbs + k*sts)*n
  const double* A; int64_t A_bs;
  const double* Q; int64_t Q_bs;
  const double* lam; int64_t lam_bs;
  const double* dt; int64_t dt_bs;
  const double* Pinf; int64_t Pinf_bs;
  const double* m0; int64_t m0_bs;
  const double* P0; int64_t P0_bs;
  const double* H; int64_t H_bs;
  const double* Y;
  const double* R; int64_t R_bs, R_ts;
  double jitter;
  double* mf; double* Pf; double* lml; double* lml_k;
  // ---- chunk mode (parallel-in-time, physs_pscan.cu): virtual series v = (b, c), c-th chunk of
  // `chunk_len` steps of series b.  fixup != 0: chunk c >= 1 restarts from the snapshot
  // (bnd_m, bnd_P)[v] of the previous chunk's last filtered state, rewrites its outputs and stops once
  // they agree with what is already stored to `delta` (relative) for `patience` consecutive steps;
  // a chunk that reaches its end without converging raises *unconverged.
  // One launch covers chunks [chunk_first, chunk_first + chunk_count) of every series; all chunks of
  // a launch have the same length (the ragged tail chunk gets its own launch).  nchunk = chunks per
  // series in total (boundary snapshots are indexed b * nchunk + c); nchunk == 0 means plain mode.
  int64_t nchunk, chunk_len, chunk_first, chunk_count;
  int from_bnd;          // chunk starts from the boundary snapshot (bnd_m, bnd_P)[v] instead of (m0, P0)
  int64_t warm;          // speculative mode (from_bnd == 0): chunk c > 0 starts `warm` steps early from (m0, P0)
  int fixup, patience;
  double delta;
  const double* bnd_m; const double* bnd_P;
  int* unconverged;
  // fix-up passes that change nothing are fixed points: a chunk whose recomputed step disagrees with the stored one sets
  // *pass_changed; a pass whose predecessor left *prev_changed == 0 returns at once (register kernels, d <= 4)
  int* pass_changed; const int* prev_changed;
  // packed hand-over (physs_kf_filter_smooth_packed_f64; register kernels, d <= 4, time-major steps, plain
  // mode): when set, the filtered moments go to pk as rows of PackedRow<D>::N doubles [m | upper triangle of P]
  // instead of (mf, Pf) -- the update leaves P bitwise symmetric, so nothing is lost
  double* pk;
};

// Outputs of the filter's reverse pass (physs_vjp.cu); any pointer may be NULL except the ones the mode needs.
struct VjpOut {
  const double* g_lml;            // [B] cotangent of lml (NULL = 1)
  double* gA; double* gQ;         // DISC_GIVEN: per-step [.., d, d] in the step layout
  double* glam; double* gPinf;    // DISC_MATERN: [B, nblk], [B, d, d]
  double* gH;                     // [B, m, d]
  double* gR_step;                // per-step [.., m, m] in the step layout, or NULL
  double* gR_sum;                 // [B, m, m] sum over the steps, or NULL
  double* gm0; double* gP0;       // [B, d], [B, d, d]
};

struct SeqSmoothArgs {
  int64_t B, T;
  int64_t sbs, sts;   // step strides (see SeqFilterArgs)
  const double* A; int64_t A_bs;
  const double* Q; int64_t Q_bs;
  const double* lam; int64_t lam_bs;
  const double* dt; int64_t dt_bs;
  const double* Pinf; int64_t Pinf_bs;
  const double* mf; const double* Pf;
  const double* Hout;
  double jitter;
  double* ms; double* Ps;
  // ---- chunk mode: chunk c of series b starts (backwards) from the carried smoothed state
  // (bnd_m, bnd_P)[v] that belongs to the first step of chunk c + 1 and treats ALL its steps as RTS
  // steps (full_state output only).  fixup as for the filter (chunks c < nchunk - 1).
  // One launch covers chunks [chunk_first, chunk_first + chunk_count) of every series; all chunks of
  // a launch have the same length (the ragged tail chunk gets its own launch).  nchunk = chunks per
  // series in total (boundary snapshots are indexed b * nchunk + c); nchunk == 0 means plain mode.
  int64_t nchunk, chunk_len, chunk_first, chunk_count;
  int carry_last;        // the last chunk also starts from its boundary snapshot (time-sharded ranges)
  int64_t warm;          // speculative mode (> 0): chunk starts `warm` steps late from the FILTERED state there
  int fixup, patience;
  double delta;
  const double* bnd_m; const double* bnd_P;
  int* unconverged;
  // host-side query (physs_kf_wave_series): when set, the launcher writes the number of series one full wave
  // of its kernel keeps resident on the device and returns WITHOUT launching
  int64_t* wave_out;
  const double* pk;      // packed filtered moments (see SeqFilterArgs::pk) read INSTEAD of (mf, Pf)
};

// physs_seq.cu: one thread per series, registers (d in {1,2,3,4,6,8})
bool seq_supported(int d, int m, int disc_mode, int nblk);
int seq_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
               const SeqFilterArgs& a);
int seq_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a);
int seq_filter_summary(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
                       const SeqFilterArgs& a, double* elems);
int seq_smooth_summary(cudaStream_t st, int d, int disc_mode, int nblk, const SeqSmoothArgs& a, double* elems);
// size dispatch shared by the plain and the chunked entry points (physs_api.cu)
bool run_filter_is_seq(int d, int m, int disc_mode, int nblk, int64_t B, int64_t nchunk);
int run_filter_any(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
                   const SeqFilterArgs& a);
int run_smooth_any(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a);

struct CviArgs {
  int64_t N;
  const double* Yt; const double* Vt;       // sites  [N,D], [N,D,D]
  const double* qm; const double* qS;       // posterior marginals of the site blocks
  const double* y;                          // data [N,P], NaN = missing
  const double* W;                          // [P,D] shared, NULL = identity
  const double* noise; int64_t noise_stride;  // Gaussian noise [.,P,P]
  double lik_param; int K; const double* ghx; const double* ghw;
  double log_param; const double* logfact;  // Poisson: log(binsize) and the device table of log-factorials (or NULL)
  const double* dm_in; const double* dS_in;
  double beta, ngj;
  int prec;                                 // sites carry (Y~, PRECISION): the 'NG_Precision' re-entry
  double* Yn; double* Vn;
  double* ell; double* dm_out; double* dS_out;
};

// physs_cvi.cu: one thread per site block (D <= 4)
bool cvi_reg_supported(int D, int P);
int cvi_reg_run(cudaStream_t st, int D, int P, int lik, bool update, const CviArgs& a);
// physs_cvi_grp.cu: one lane group per site block (D > 4)
int cvi_grp_run(cudaStream_t st, int D, int P, int lik, bool update, const CviArgs& a);

// physs_grp.cu: one lane group per series, shared-memory resident, runtime (d, m)
bool grp_supported(int d, int m);
int grp_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
               const SeqFilterArgs& a);
int grp_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a);

// physs_rt.cu: one lane group per series, register-tiled products, compile-time padded dims (d <= 32)
bool rt_supported(int d, int m);
bool vjp_supported(int d, int m, int disc_mode, int nblk);
int kf_vjp(cudaStream_t st, int d, int m, int disc_mode, int nblk, const SeqFilterArgs& a, const VjpOut& o);
// physs_vjp_grp.cu: the same reverse pass for general (d <= 32, m <= d), DISC_GIVEN, one lane group per series
bool grp_vjp_supported(int d, int m, int disc_mode);
int grp_kf_vjp(cudaStream_t st, int d, int m, bool h_identity, const SeqFilterArgs& a, const VjpOut& o);
int rt_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity, const SeqFilterArgs& a);
int rt_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a);
int rt_filter_summary(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, const SeqFilterArgs& a,
                      int64_t cfirst, int64_t ccount, double* elems);
int rt_smooth_summary(cudaStream_t st, int d, int disc_mode, int nblk, const SeqSmoothArgs& a, double* elems);

// physs_colloc.cu: collocation (EKF) filter step, d <= 4 (descriptor arrays are HOST pointers)
int colloc_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, const SeqFilterArgs& a,
                  int pc, const double* res_w, int n_terms, const int32_t* t_out, const int32_t* t_kind,
                  const int32_t* t_idx, const double* t_coef, const double* forcing, const double* y_pseudo,
                  const double* boundary, int observe_data);

// physs_spd.cu: batched inverse of small SPD matrices (precision sites -> covariance sites)
int spd_inverse(cudaStream_t st, int64_t N, int D, const double* A, double jitter, double* out);

// physs_pscan.cu: parallel-in-time chunked associative scan
int64_t pscan_workspace_doubles(int64_t B, int64_t T, int d, int64_t chunk_len);
int pscan_filter_local(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, SeqFilterArgs a,
                       int64_t chunk_len, double* ws, double* total_out);
// physs_pscan.cu: out[b] = sum_k (x[b, k] - sub[b, k]) over a [B, T] array with strides (sbs, sts); sub may be NULL
int sum_steps(cudaStream_t st, int64_t B, int64_t T, int64_t sbs, int64_t sts, const double* x, const double* sub,
              double* scratch, double* out);
int pscan_filter_finish(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, SeqFilterArgs a,
                        int64_t chunk_len, double* ws, bool had_total, const double* start_m,
                        const double* start_P, int polish, double delta, int patience, int* status_out);
int pscan_filter_fold(cudaStream_t st, int d, int64_t B, int64_t K, const double* totals, const double* m0,
                      int64_t m0_bs, const double* P0, int64_t P0_bs, double* m_out, double* P_out);
int pscan_smooth_local(cudaStream_t st, int d, int disc_mode, int nblk, SeqSmoothArgs a, int64_t chunk_len,
                       double* ws, double* total_out);
int pscan_smooth_finish(cudaStream_t st, int d, int mo, int disc_mode, int nblk, SeqSmoothArgs a, int64_t chunk_len,
                        double* ws, const double* start_m, const double* start_P);
int pscan_filter_spec(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, SeqFilterArgs a,
                      int64_t chunk_len, int64_t warm, int polish, double delta, int patience, double* ws,
                      int* status_out);
int pscan_smooth_spec(cudaStream_t st, int d, int mo, int disc_mode, int nblk, SeqSmoothArgs a, int64_t chunk_len,
                      int64_t warm, int polish, double delta, int patience, double* ws, int* status_out);
int pscan_smooth_fold(cudaStream_t st, int d, int64_t B, int64_t K, const double* totals, const double* m0,
                      const double* P0, double* m_out, double* P_out);

}  // namespace physs
