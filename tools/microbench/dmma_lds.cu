// dmma_lds.cu -- what does a DMMA.8x8x4 k-loop fed from shared memory cost per instruction?  (measurement aid for
// csrc/physs_kron.cu: the tile GEMM measured ~30 cycles per DMMA per scheduler against 16 in fp64_pipes.cu)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_lds dmma_lds.cu ; run: ./dmma_lds
#include <cuda_runtime.h>
#include <stdio.h>
constexpr int LDA = 132, LDB = 36, KC = 128;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MT, int MODE>
__global__ void k(double* out, int iters, long long* cyc) {
  extern __shared__ double sm[];
  double* As = sm; double* Bs = sm + 40 * LDA;
  for (int i = threadIdx.x; i < 40 * LDA + KC * LDB; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  const int strip = warp & 3;
  double acc[MT][2];
  for (int m = 0; m < MT; ++m) acc[m][0] = acc[m][1] = 0.0;
  const double* as = As + gq * LDA + tq;
  const double* bs = Bs + tq * LDB + 8 * strip + gq;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // loads right before use
      for (int kt = 0; kt < 32; ++kt) {
        const double b = bs[kt * 4 * LDB];
#pragma unroll
        for (int m = 0; m < MT; ++m) dmma(acc[m][0], acc[m][1], as[8 * m * LDA + 4 * kt], b);
      }
    } else if (MODE == 1) {   // software pipelined (as in gemm_tiles)
      double a0[MT], a1[MT], b0, b1;
      b0 = bs[0];
#pragma unroll
      for (int m = 0; m < MT; ++m) a0[m] = as[8 * m * LDA];
      for (int kt = 0; kt + 2 <= 32; kt += 2) {
        b1 = bs[(kt + 1) * 4 * LDB];
#pragma unroll
        for (int m = 0; m < MT; ++m) a1[m] = as[8 * m * LDA + 4 * kt + 4];
#pragma unroll
        for (int m = 0; m < MT; ++m) dmma(acc[m][0], acc[m][1], a0[m], b0);
        if (kt + 2 < 32) {
          b0 = bs[(kt + 2) * 4 * LDB];
#pragma unroll
          for (int m = 0; m < MT; ++m) a0[m] = as[8 * m * LDA + 4 * kt + 8];
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) dmma(acc[m][0], acc[m][1], a1[m], b1);
      }
    } else {                  // registers only
      double a[MT], b = bs[0];
#pragma unroll
      for (int m = 0; m < MT; ++m) a[m] = as[8 * m * LDA];
      for (int kt = 0; kt < 32; ++kt) {
#pragma unroll
        for (int m = 0; m < MT; ++m) dmma(acc[m][0], acc[m][1], a[m], b);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int m = 0; m < MT; ++m) s += acc[m][0] + acc[m][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MT, int MODE>
void run(int warps, const char* name) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const size_t smem = (40 * LDA + KC * LDB) * 8;
  cudaFuncSetAttribute(k<MT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 200;
  k<MT, MODE><<<148, warps * 32, smem>>>(out, iters, cyc);
  k<MT, MODE><<<148, warps * 32, smem>>>(out, iters, cyc);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  const double per_sched = (double)c / (iters * 32.0 * MT * (warps / 4.0));
  printf("{\"variant\": \"%s\", \"mt\": %d, \"warps\": %d, \"cycles_per_dmma_per_scheduler\": %.2f, \"err\": \"%s\"}\n", name, MT, warps,
         per_sched, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<5, 2>(4, "registers"); run<5, 0>(4, "lds_just_in_time"); run<5, 1>(4, "lds_pipelined");
  run<5, 2>(8, "registers"); run<5, 0>(8, "lds_just_in_time"); run<5, 1>(8, "lds_pipelined");
  run<4, 1>(4, "lds_pipelined"); run<1, 1>(4, "lds_pipelined"); run<1, 1>(8, "lds_pipelined");
  return 0;
}
