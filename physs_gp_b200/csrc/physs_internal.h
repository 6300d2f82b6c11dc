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
};

struct SeqSmoothArgs {
  int64_t B, T;
  const double* A; int64_t A_bs;
  const double* Q; int64_t Q_bs;
  const double* lam; int64_t lam_bs;
  const double* dt; int64_t dt_bs;
  const double* Pinf; int64_t Pinf_bs;
  const double* mf; const double* Pf;
  const double* Hout;
  double jitter;
  double* ms; double* Ps;
};

// physs_seq.cu: one thread per series, registers (d in {1,2,3,4,6,8})
bool seq_supported(int d, int m, int disc_mode, int nblk);
int seq_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
               const SeqFilterArgs& a);
int seq_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a);

// physs_grp.cu: one lane group per series, shared-memory resident, runtime (d, m)
bool grp_supported(int d, int m);
int grp_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
               const SeqFilterArgs& a);
int grp_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a);

}  // namespace physs
