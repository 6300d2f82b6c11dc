// physs_spd.cu -- batched inverse of small SPD matrices, one thread per matrix (D <= 8).
//
// Used for precision-parameterised sites: the reference's sequential filter raises for them
// (kf_update_step_with_lik_precision, computation/filters/kalman_filter.py:43-126, is dead code) and only its
// parallel filter accepts R_inv (parallel_kalman_filter.py:34-71,117-141).  The b200 backends take R_inv by turning
// it into R = chol_solve(chol(R_inv), I) on the device -- the same factor-and-solve the reference's own precision
// elements apply to it -- and then run the ordinary covariance-form recursion.
#include "physs_core.cuh"
#include "physs_internal.h"

namespace physs {

template <int D>
__global__ void __launch_bounds__(128) spd_inverse_kernel(int64_t N, const double* __restrict__ A, double jitter,
                                                          double* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double M[D][D], L[D][D], rd[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) M[i][j] = A[n * D * D + i * D + j] + (i == j ? jitter : 0.0);
  }
  chol_lower<D>(M, L, rd);
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = (i == c) ? 1.0 : 0.0;
    chol_solve_vec<D>(L, rd, x);
#pragma unroll
    for (int i = 0; i < D; ++i) out[n * D * D + i * D + c] = x[i];
  }
}

int spd_inverse(cudaStream_t st, int64_t N, int D, const double* A, double jitter, double* out) {
  if (N <= 0) return PHYSS_OK;
  const unsigned grid = (unsigned)((N + 127) / 128);
  switch (D) {
#define CASE(D_) case D_: spd_inverse_kernel<D_><<<grid, 128, 0, st>>>(N, A, jitter, out); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: return set_error(PHYSS_ERR_UNSUPPORTED, "spd inverse: D <= 8");
  }
  return cuda_status(cudaGetLastError(), "spd_inverse_kernel launch");
}

}  // namespace physs
