// fp64_pipes.cu -- micro-benchmark behind the "DMMA or DFMA?" decision.
// Measures on the box it runs on, per SM sub-partition (one warp scheduler):
//   * issue interval and dependent-issue latency of DFMA and of DMMA.8x8x4 (mma.sync.m8n8k4.f64),
//   * aggregate FP64 throughput of each with 1 / 2 / 4 / 8 warps per scheduler and ILP 1 .. 8,
//   * whether the two pipes overlap (half the warps DFMA, half DMMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu ; run: ./fp64_pipes
#include <cuda_runtime.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

template <int ILP>
__device__ __forceinline__ double run_dfma(double x, double y) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], y, x);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  return s;
}

template <int ILP>
__device__ __forceinline__ double run_dmma(double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  return s;
}

// mode 0: all DFMA, 1: all DMMA, 2: even warps DFMA / odd warps DMMA
template <int ILP>
__global__ void probe(int mode, double* out, long long* cyc, double x, double y) {
  const int warp = threadIdx.x >> 5;
  const bool dmma = (mode == 1) || (mode == 2 && (warp & 1));
  __syncthreads();
  const long long t0 = clock64();
  double r = dmma ? run_dmma<ILP>(x, y) : run_dfma<ILP>(x, y);
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  if (r == 123.456) out[0] = r;
}

template <int ILP>
int one(int mode, int warps_per_sm, int sms, double* out, long long* cyc, long long* hcyc) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  probe<ILP><<<sms, 32 * warps_per_sm>>>(mode, out, cyc, 1.0, 0.999);      // warm-up
  CHECK(cudaEventRecord(e0));
  probe<ILP><<<sms, 32 * warps_per_sm>>>(mode, out, cyc, 1.0, 0.999);
  CHECK(cudaEventRecord(e1));
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  CHECK(cudaEventElapsedTime(&ms, e0, e1));
  CHECK(cudaMemcpy(hcyc, cyc, sizeof(long long) * sms * warps_per_sm, cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < sms * warps_per_sm; ++i) mean += (double)hcyc[i];
  mean /= sms * warps_per_sm;
  const double per_instr = mean / ((double)ITERS * ILP);             // cycles per instruction per warp
  // flop: DFMA warp-instruction = 64, DMMA.8x8x4 = 512
  double n_dfma = 0, n_dmma = 0;
  const double per_warp = (double)ITERS * ILP;
  if (mode == 0) n_dfma = per_warp * warps_per_sm * sms;
  else if (mode == 1) n_dmma = per_warp * warps_per_sm * sms;
  else { n_dfma = per_warp * (warps_per_sm - warps_per_sm / 2) * sms; n_dmma = per_warp * (warps_per_sm / 2) * sms; }
  const double tflops = (n_dfma * 64 + n_dmma * 512) / (ms * 1e-3) / 1e12;
  printf("{\"mode\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"cycles_per_instr_per_warp\": %.2f, \"ms\": %.4f, \"tflops\": %.2f}\n",
         mode == 0 ? "dfma" : mode == 1 ? "dmma" : "mixed", ILP, warps_per_sm, per_instr, ms, tflops);
  return 0;
}

int main() {
  int sms = 0;
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; long long* cyc;
  CHECK(cudaMalloc(&out, 8)); CHECK(cudaMalloc(&cyc, sizeof(long long) * sms * 64));
  long long* hcyc = (long long*)malloc(sizeof(long long) * sms * 64);
  const int wps[] = {4, 8, 16, 32};
  for (int mode = 0; mode < 3; ++mode)
    for (int w : wps) {
      if (one<1>(mode, w, sms, out, cyc, hcyc)) return 1;
      if (one<2>(mode, w, sms, out, cyc, hcyc)) return 1;
      if (one<4>(mode, w, sms, out, cyc, hcyc)) return 1;
      if (one<8>(mode, w, sms, out, cyc, hcyc)) return 1;
    }
  return 0;
}
