// physs_cvi.cu -- CVI natural-gradient site update and expected log-likelihoods, one thread per site
// block (block size D <= 4, registers).  Every (series, time step) block is independent, so this is a
// flat, fully coalesced pass over N = B*T blocks: HBM-bound, 8*(3 D^2 + 3 D + P) bytes per block for
// the fused update (sites + posterior marginals in, sites out).
//
// Replaces natural_gradients(FullConjugateGaussian) + cvi_block_update + theta<->lambda
// (cvi_nat_grad.py:346-410,47-87; exponential_family_transforms.py:25-83) and the block expected
// log-likelihood (expected_log_likelihoods.py:90-117; elbos.py:163-194).
#include "physs_cvi_core.cuh"
#include "physs_internal.h"

namespace physs {

template <int N>
__device__ __forceinline__ void ldv(const double* __restrict__ src, double (&dst)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) dst[i] = src[i];
}
template <int N>
__device__ __forceinline__ void stv(double* __restrict__ dst, const double (&src)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) dst[i] = src[i];
}

template <int D, int P, int LIK, bool UPDATE>
__global__ void __launch_bounds__(128) cvi_site_kernel(const CviArgs p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= p.N) return;
  double qm[D], qS[D][D], dm[D], dS[D][D];
  ldv<D>(p.qm + n * D, qm);
  ldv<D * D>(p.qS + n * D * D, *reinterpret_cast<double (*)[D * D]>(&qS[0][0]));
  // the sites are loaded up front: behind the ELL (and its optional stores, which the compiler must assume alias) their
  // latency was a second, fully exposed round trip to HBM (ncu: 9.8 long-scoreboard stalls per issue)
  double Yt[D], Vt[D][D];
  if (UPDATE) {
    ldv<D>(p.Yt + n * D, Yt);
    ldv<D * D>(p.Vt + n * D * D, *reinterpret_cast<double (*)[D * D]>(&Vt[0][0]));
  }
  double ell = 0.0;
  if (LIK == CVI_LIK_GIVEN) {
    ldv<D>(p.dm_in + n * D, dm);
    ldv<D * D>(p.dS_in + n * D * D, *reinterpret_cast<double (*)[D * D]>(&dS[0][0]));
  } else {
    double y[P], W[P][D], noise[P][P];
    ldv<P>(p.y + n * P, y);
#pragma unroll
    for (int a = 0; a < P; ++a) {
#pragma unroll
      for (int k = 0; k < D; ++k) W[a][k] = p.W ? p.W[a * D + k] : (a == k ? 1.0 : 0.0);
    }
    if (LIK == CVI_LIK_GAUSS)
      ldv<P * P>(p.noise + n * p.noise_stride, *reinterpret_cast<double (*)[P * P]>(&noise[0][0]));
    ell = cvi_ell_grads<D, P, LIK>(qm, qS, y, W, noise, p.lik_param, p.K, p.ghx, p.ghw, dm, dS, p.ell != nullptr,
                                   p.log_param, p.logfact, true);
    if (p.ell) p.ell[n] = ell;
    if (p.dm_out) stv<D>(p.dm_out + n * D, dm);
    if (p.dS_out) stv<D * D>(p.dS_out + n * D * D, *reinterpret_cast<double (*)[D * D]>(&dS[0][0]));
  }
  if (UPDATE) {
    double Yn[D], Vn[D][D];
    cvi_site_update<D>(Yt, Vt, qm, qS, dm, dS, p.beta, p.ngj, Yn, Vn, p.prec != 0);
    stv<D>(p.Yn + n * D, Yn);
    stv<D * D>(p.Vn + n * D * D, *reinterpret_cast<double (*)[D * D]>(&Vn[0][0]));
  }
}

// log(y!) for the counts 0..255, evaluated once per process on the host (lgammal, rounded to double) and kept in device
// memory.  The first request that arrives while its stream is being captured into a CUDA graph gets no table (the
// kernels then call lgamma themselves): the upload is a synchronous copy, which a capture does not allow.
__device__ double g_logfact[kLogFactN];
static const double* logfact_table(cudaStream_t st) {
  static const double* table = nullptr;
  static int device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (table && dev == device) return table;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;
  double h[kLogFactN];
  for (int i = 0; i < kLogFactN; ++i) h[i] = (double)lgammal((long double)i + 1.0L);
  void* ptr = nullptr;
  if (cudaMemcpyToSymbol(g_logfact, h, sizeof(h)) != cudaSuccess || cudaGetSymbolAddress(&ptr, g_logfact) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  table = static_cast<const double*>(ptr);
  device = dev;
  return table;
}

template <int D, int P, int LIK, bool UPDATE>
static int launch(cudaStream_t st, const CviArgs& a0) {
  CviArgs a = a0;
  if (LIK == CVI_LIK_POISSON_EXP) {
    a.log_param = log(a.lik_param);
    a.logfact = logfact_table(st);
  }
  const int block = 128;
  const int64_t grid = (a.N + block - 1) / block;
  cvi_site_kernel<D, P, LIK, UPDATE><<<(unsigned)grid, block, 0, st>>>(a);
  return cuda_status(cudaGetLastError(), "cvi_site_kernel launch");
}

template <int D, int P, bool UPDATE>
static int by_lik(cudaStream_t st, const CviArgs& a, int lik) {
  switch (lik) {
    case CVI_LIK_GAUSS: return launch<D, P, CVI_LIK_GAUSS, UPDATE>(st, a);
    case CVI_LIK_POISSON_EXP: return launch<D, P, CVI_LIK_POISSON_EXP, UPDATE>(st, a);
    case CVI_LIK_BERNOULLI_PROBIT: return launch<D, P, CVI_LIK_BERNOULLI_PROBIT, UPDATE>(st, a);
    case CVI_LIK_GIVEN:
      if (UPDATE) return launch<D, 1, CVI_LIK_GIVEN, UPDATE>(st, a);
      break;
  }
  return set_error(PHYSS_ERR_BAD_ARG, "cvi: unknown likelihood kind");
}

template <int D, bool UPDATE>
static int by_p(cudaStream_t st, const CviArgs& a, int P, int lik) {
  if (lik == CVI_LIK_GIVEN) return by_lik<D, 1, UPDATE>(st, a, lik);
  if (P == 1) return by_lik<D, 1, UPDATE>(st, a, lik);
  if (D >= 2 && P == 2) return by_lik<D, (D >= 2 ? 2 : 1), UPDATE>(st, a, lik);
  if (D >= 3 && P == 3) return by_lik<D, (D >= 3 ? 3 : 1), UPDATE>(st, a, lik);
  if (D >= 4 && P == 4) return by_lik<D, (D >= 4 ? 4 : 1), UPDATE>(st, a, lik);
  return set_error(PHYSS_ERR_UNSUPPORTED, "cvi: need 1 <= P <= D");
}

bool cvi_reg_supported(int D, int P) { return D >= 1 && D <= 4 && P >= 1 && P <= D; }

int cvi_reg_run(cudaStream_t st, int D, int P, int lik, bool update, const CviArgs& a) {
  if (!cvi_reg_supported(D, lik == CVI_LIK_GIVEN ? 1 : P))
    return set_error(PHYSS_ERR_UNSUPPORTED, "cvi: register path needs D <= 4");
#define X(D_) \
  if (D == D_) return update ? by_p<D_, true>(st, a, P, lik) : by_p<D_, false>(st, a, P, lik);
  X(1) X(2) X(3) X(4)
#undef X
  return set_error(PHYSS_ERR_UNSUPPORTED, "cvi: unreachable");
}

}  // namespace physs
