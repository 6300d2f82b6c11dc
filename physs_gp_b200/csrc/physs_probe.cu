// physs_probe.cu -- FP64 FMA throughput probe.  MEASURED_PEAKS.json carries HBM and bf16 peaks only; the
// FP64-pipe roofline of the d >~ 10 regime needs the DFMA rate of THIS GPU, measured
// the same way the kernels use the pipe: register-resident dependent-chain-free fused multiply-adds.
#include "physs_internal.h"

namespace physs {

__global__ void __launch_bounds__(256) fp64_probe_kernel(int64_t iters, double seed, double* __restrict__ out) {
  // 8 independent accumulator chains per thread keep the FP64 pipe full
  double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
  double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double x = 0.999999, y = 1e-9;
  for (int64_t i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) out[0] = s;     // never true; keeps the loop alive
}

}  // namespace physs

extern "C" {

/* Launches `blocks` x 256 threads, each doing iters * 8 FMAs (2 flop each).  Time it with CUDA events on
 * `stream`; flops = blocks * 256 * iters * 16.  out: one device double (never written in practice). */
int physs_fp64_probe(void* stream, int32_t blocks, int64_t iters, double* out) {
  if (blocks < 1 || iters < 1 || !out) return physs::set_error(PHYSS_ERR_BAD_ARG, "fp64 probe: bad arguments");
  physs::fp64_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, out);
  return physs::cuda_status(cudaGetLastError(), "fp64_probe_kernel launch");
}

}
