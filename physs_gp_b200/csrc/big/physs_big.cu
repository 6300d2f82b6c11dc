// physs_big.cu -- large-block sequential Kalman filter / RTS smoother (one series, state dim d up to a few
// thousand): BASELINE config 2 (spatio-temporal separable Matern-3/2 x RBF, 200 spatial points -> d = 400,
// m = 200, T = 5000).  libphyss_b200_big.so, C ABI in include/physs_b200_big.h.
//
// At this size one step IS dense linear algebra on 400 x 400 fp64 blocks, so -- as the task's rules allow for
// plain library GEMMs -- the matrix products, Cholesky factorisations and triangular solves are cuBLAS /
// cuSOLVER calls enqueued on the caller's stream; the glue that is specific to the reference's semantics
// (missing-data masks on H P H^T, jitter placement, mask-to-identity for the lml, log-determinant /
// Mahalanobis accumulation) is hand-written.  Same reference functions as the small-block kernels:
// kalman_filter.py:144-241,439-485 and rts_smoother.py:48-106,162-192.
//
// Row-major <-> column-major: every matrix here is row-major; cuBLAS / cuSOLVER see its transpose.  Symmetric
// matrices (P, S, Q) are the same in both views; products use C^T = B^T A^T; the row-major d x m buffer P H^T
// is, column-major, the m x d right-hand side (P H^T)^T of the gain solve, so potrs leaves K (row-major d x m)
// in place.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/physs_b200_big.h"

namespace {

thread_local char g_err[256] = "";
int fail(int code, const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); return code; }

struct Handles {
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t solver = nullptr;
};
// one pair of handles per host thread and device (handles are cheap to keep, expensive to create)
Handles* handles() {
  thread_local Handles h[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!h[dev].blas) {
    if (cublasCreate(&h[dev].blas) != CUBLAS_STATUS_SUCCESS) return nullptr;
    if (cusolverDnCreate(&h[dev].solver) != CUSOLVER_STATUS_SUCCESS) return nullptr;
  }
  return &h[dev];
}

#define CB(x) do { if ((x) != CUBLAS_STATUS_SUCCESS) return fail(PHYSS_BIG_ERR_LIB, "cuBLAS call failed: " #x); } while (0)
#define CS(x) do { if ((x) != CUSOLVER_STATUS_SUCCESS) return fail(PHYSS_BIG_ERR_LIB, "cuSOLVER call failed: " #x); } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s: %s", #x, cudaGetErrorString(e_)); return PHYSS_BIG_ERR_CUDA; } } while (0)

// row-major C[n x m] = alpha * opA(A) opB(B) + beta * C, with A stored [n x k] (or [k x n] if ta), B [k x m] (or [m x k])
cublasStatus_t gemm_rm(cublasHandle_t h, bool ta, bool tb, int n, int m, int k, double alpha, const double* A, int lda,
                       const double* B, int ldb, double beta, double* C, int ldc) {
  // column-major view: C^T [m x n] = opB(B)^T opA(A)^T; a row-major X is the column-major X^T
  return cublasDgemm(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, m, n, k, &alpha, B, ldb, A, lda,
                     &beta, C, ldc);
}
// row-major y[n] = alpha * opA(A) x + beta * y, A stored [n x k] (ta: [k x n])
cublasStatus_t gemv_rm(cublasHandle_t h, bool ta, int n, int k, double alpha, const double* A, int lda, const double* x,
                       double beta, double* y) {
  // column-major view of A is A^T with dims (cols, rows) = (lda-ordered)
  if (!ta) return cublasDgemv(h, CUBLAS_OP_T, k, n, &alpha, A, lda, x, 1, &beta, y, 1);
  return cublasDgemv(h, CUBLAS_OP_N, n, k, &alpha, A, lda, x, 1, &beta, y, 1);
}

// ------------------------------------------------------------------------------------ glue kernels
// PHt[i][a] = 0 for missing a ; v[a] = obs ? y[a] - mu[a] : 0 ; w = v
__global__ void mask_gain_rhs(double* __restrict__ PHt, int d, int m, const double* __restrict__ y,
                              const double* __restrict__ mu, double* __restrict__ v, double* __restrict__ w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < d * m) {
    const int a = idx % m;
    if (y[a] != y[a]) PHt[idx] = 0.0;
  }
  if (idx < m) {
    const double ya = y[idx];
    const double r = (ya != ya) ? 0.0 : ya - mu[idx];
    v[idx] = r;
    w[idx] = r;
  }
}
// S = M (H P H^T) M + R (R not masked) ; Sj = S + jitter I ; Sm = mask-to-identity(S)
__global__ void build_S(const double* __restrict__ HPHt, const double* __restrict__ R, const double* __restrict__ y,
                        int m, double jitter, double* __restrict__ S, double* __restrict__ Sj, double* __restrict__ Sm) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int a = idx / m, c = idx % m;
  const bool oa = !(y[a] != y[a]), oc = !(y[c] != y[c]);
  const double s = ((oa && oc) ? HPHt[idx] : 0.0) + R[idx];
  S[idx] = s;
  Sj[idx] = s + (a == c ? jitter : 0.0);
  Sm[idx] = (oa && oc) ? s : (a == c ? 1.0 : 0.0);
}
// lml += -1/2 (nobs log 2pi + 2 sum log diag(L) + v . w)   (single block)
__global__ void lml_accumulate(const double* __restrict__ L, int m, const double* __restrict__ v,
                               const double* __restrict__ w, const double* __restrict__ y, double* __restrict__ lml) {
  __shared__ double red[3][32];
  double ld = 0.0, mh = 0.0, no = 0.0;
  for (int a = threadIdx.x; a < m; a += blockDim.x) {
    ld += log(L[(size_t)a * m + a]);
    mh += v[a] * w[a];
    no += (y[a] != y[a]) ? 0.0 : 1.0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ld += __shfl_xor_sync(0xffffffffu, ld, o);
    mh += __shfl_xor_sync(0xffffffffu, mh, o);
    no += __shfl_xor_sync(0xffffffffu, no, o);
  }
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
  if (ln == 0) { red[0][wid] = ld; red[1][wid] = mh; red[2][wid] = no; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
    lml[0] += -0.5 * (c * 1.8378770664093454835606594728112 + 2.0 * a + b);
  }
}
__global__ void axpby_kernel(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b,
                             double beta, int64_t n, int diag_stride, double diag_add) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = a[i] + beta * b[i];
  if (diag_stride > 0 && i % (diag_stride + 1) == 0) v += diag_add;
  out[i] = v;
}
__global__ void add_diag(double* __restrict__ A, int n, double x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(size_t)i * n + i] += x;
}
inline unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }
// cusolverDnDpotrf reports a non-PD input only through *info and leaves a finite, partially factored matrix
// behind.  The reference's jnp.linalg.cholesky gives NaN there and the header promises NaN on numerical
// failure, so (stream-ordered, no host sync) a failed factorisation poisons the factor -- every later potrs
// then yields NaN -- and, when given, the running lml.
__global__ void poison_on_potrf_failure(const int* __restrict__ info, double* __restrict__ factor,
                                        double* __restrict__ lml) {
  if (*info != 0) {
    factor[0] = nan("");
    if (lml) lml[0] = nan("");
  }
}

struct Ws {
  double *W, *PHt, *HPHt, *S, *Sj, *Sm, *KS, *mp, *mu, *v, *w, *lwork, *dP, *Pp, *T1, *ms, *Ps0, *Ps1, *HPs;
  int* info;
  int lwork_n;
};

int64_t ws_doubles(int d, int m, int lwork) {
  const int64_t dd = (int64_t)d * d, dm = (int64_t)d * m, mm = (int64_t)m * m;
  return 6 * dd + 2 * dm + 4 * mm + 5 * (int64_t)(d > m ? d : m) + lwork + 16;
}
Ws carve(double* p, int d, int m, int lwork) {
  const int64_t dd = (int64_t)d * d, dm = (int64_t)d * m, mm = (int64_t)m * m;
  const int64_t vec = d > m ? d : m;
  Ws w{};
  w.W = p; p += dd; w.dP = p; p += dd; w.Pp = p; p += dd; w.T1 = p; p += dd; w.Ps0 = p; p += dd; w.Ps1 = p; p += dd;
  w.PHt = p; p += dm; w.KS = p; p += dm;
  w.HPHt = p; p += mm; w.S = p; p += mm; w.Sj = p; p += mm; w.Sm = p; p += mm;
  w.mp = p; p += vec; w.mu = p; p += vec; w.v = p; p += vec; w.w = p; p += vec; w.ms = p; p += vec;
  w.lwork = p; p += lwork; w.lwork_n = lwork;
  w.info = reinterpret_cast<int*>(p);
  w.HPs = w.KS;     // smoother projection scratch [mo x d] shares the [d x m] slot (mo <= d enforced)
  return w;
}

int potrf_lwork(cusolverDnHandle_t s, int n, double* A, int* lwork) {
  CS(cusolverDnDpotrf_bufferSize(s, CUBLAS_FILL_MODE_LOWER, n, A, n, lwork));
  return 0;
}

}  // namespace

extern "C" {

const char* physs_big_last_error(void) { return g_err; }

int64_t physs_big_workspace_bytes(int32_t d, int32_t m) {
  if (d < 1 || m < 1) return 0;
  Handles* h = handles();
  if (!h) return 0;
  int l1 = 0, l2 = 0;
  if (cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, d, nullptr, d, &l1) != CUSOLVER_STATUS_SUCCESS) return 0;
  if (cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, m, nullptr, m, &l2) != CUSOLVER_STATUS_SUCCESS) return 0;
  return 8 * ws_doubles(d, d > m ? d : m, l1 > l2 ? l1 : l2);
}

int physs_kf_filter_big_f64(void* stream, int64_t T, int32_t d, int32_t m, const double* A, const double* Q,
                            const int32_t* disc_index, const double* m0, const double* P0, const double* H,
                            const double* Y, const double* R, int64_t R_tstride, double jitter, void* ws,
                            int64_t ws_bytes, double* mf, double* Pf, double* lml) {
  if (T < 1 || d < 1 || m < 1 || m > d) return fail(PHYSS_BIG_ERR_BAD_ARG, "big filter: bad sizes (need 1 <= m <= d)");
  if (!A || !Q || !m0 || !P0 || !H || !Y || !R || !ws || !mf || !Pf || !lml)
    return fail(PHYSS_BIG_ERR_BAD_ARG, "big filter: null required pointer");
  Handles* h = handles();
  if (!h) return fail(PHYSS_BIG_ERR_LIB, "big filter: cuBLAS / cuSOLVER handle creation failed");
  cudaStream_t st = (cudaStream_t)stream;
  CB(cublasSetStream(h->blas, st));
  CS(cusolverDnSetStream(h->solver, st));
  int lwork = 0;
  {
    int l2 = 0;
    CS(cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, m, nullptr, m, &l2));
    CS(cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, d, nullptr, d, &lwork));
    if (l2 > lwork) lwork = l2;
  }
  if (ws_bytes < 8 * ws_doubles(d, d, lwork)) return fail(PHYSS_BIG_ERR_BAD_ARG, "big filter: workspace too small");
  Ws w = carve((double*)ws, d, d, lwork);
  const int64_t dd = (int64_t)d * d;
  CU(cudaMemsetAsync(lml, 0, sizeof(double), st));
  const double* m_prev = m0;
  const double* P_prev = P0;
  for (int64_t k = 0; k < T; ++k) {
    const int64_t di = disc_index ? disc_index[k] : k;
    const double* Ak = A + di * dd;
    const double* Qk = Q + di * dd;
    double* m_k = mf + k * d;
    double* P_k = Pf + k * dd;
    const double* y = Y + k * m;
    const double* Rk = R + k * R_tstride;
    // predict: m_ = A m ; P_ = A P A^T + Q (into the output slot)
    CB(gemv_rm(h->blas, false, d, d, 1.0, Ak, d, m_prev, 0.0, w.mp));
    CB(gemm_rm(h->blas, false, false, d, d, d, 1.0, Ak, d, P_prev, d, 0.0, w.W, d));
    CU(cudaMemcpyAsync(P_k, Qk, dd * 8, cudaMemcpyDeviceToDevice, st));
    CB(gemm_rm(h->blas, false, true, d, d, d, 1.0, w.W, d, Ak, d, 1.0, P_k, d));
    // update
    CB(gemm_rm(h->blas, false, true, d, m, d, 1.0, P_k, d, H, d, 0.0, w.PHt, m));          // P_ H^T  [d x m]
    CB(gemv_rm(h->blas, false, m, d, 1.0, H, d, w.mp, 0.0, w.mu));                          // H m_
    mask_gain_rhs<<<blocks_for((int64_t)d * m), 256, 0, st>>>(w.PHt, d, m, y, w.mu, w.v, w.w);
    CB(gemm_rm(h->blas, false, false, m, m, d, 1.0, H, d, w.PHt, m, 0.0, w.HPHt, m));       // H P_ H^T (cols masked)
    build_S<<<blocks_for((int64_t)m * m), 256, 0, st>>>(w.HPHt, Rk, y, m, jitter, w.S, w.Sj, w.Sm);
    CS(cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, m, w.Sj, m, w.lwork, w.lwork_n, w.info));
    poison_on_potrf_failure<<<1, 1, 0, st>>>(w.info, w.Sj, nullptr);
    CS(cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, m, d, w.Sj, m, w.PHt, m, w.info));  // -> K [d x m]
    CU(cudaMemcpyAsync(m_k, w.mp, (size_t)d * 8, cudaMemcpyDeviceToDevice, st));
    CB(gemv_rm(h->blas, false, d, m, 1.0, w.PHt, m, w.v, 1.0, m_k));                        // m = m_ + K v
    CB(gemm_rm(h->blas, false, false, d, m, m, 1.0, w.PHt, m, w.S, m, 0.0, w.KS, m));       // K S
    CB(gemm_rm(h->blas, false, true, d, d, m, -1.0, w.KS, m, w.PHt, m, 1.0, P_k, d));       // P = P_ - K S K^T
    // lml: un-jittered S, missing rows / cols -> identity
    CS(cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, m, w.Sm, m, w.lwork, w.lwork_n, w.info));
    poison_on_potrf_failure<<<1, 1, 0, st>>>(w.info, w.Sm, lml);
    CS(cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, m, 1, w.Sm, m, w.w, m, w.info));
    lml_accumulate<<<1, 256, 0, st>>>(w.Sm, m, w.v, w.w, y, lml);
    m_prev = m_k;
    P_prev = P_k;
  }
  CU(cudaGetLastError());
  return PHYSS_BIG_OK;
}

int physs_rts_smooth_big_f64(void* stream, int64_t T, int32_t d, const double* A, const double* Q,
                             const int32_t* disc_index, const double* mf, const double* Pf, const double* Hout,
                             int32_t mo, double jitter, void* ws, int64_t ws_bytes, double* ms, double* Ps) {
  if (T < 1 || d < 1 || mo < 0 || mo > d) return fail(PHYSS_BIG_ERR_BAD_ARG, "big smoother: bad sizes");
  if (!A || !Q || !mf || !Pf || !ws || !ms || !Ps || (mo > 0 && !Hout))
    return fail(PHYSS_BIG_ERR_BAD_ARG, "big smoother: null required pointer");
  Handles* h = handles();
  if (!h) return fail(PHYSS_BIG_ERR_LIB, "big smoother: cuBLAS / cuSOLVER handle creation failed");
  cudaStream_t st = (cudaStream_t)stream;
  CB(cublasSetStream(h->blas, st));
  CS(cusolverDnSetStream(h->solver, st));
  int lwork = 0;
  CS(cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, d, nullptr, d, &lwork));
  if (ws_bytes < 8 * ws_doubles(d, d, lwork)) return fail(PHYSS_BIG_ERR_BAD_ARG, "big smoother: workspace too small");
  Ws w = carve((double*)ws, d, d, lwork);
  const int64_t dd = (int64_t)d * d;
  const int mp_ = mo == 0 ? d : mo;
  // full-state recursion lives in (ms_cur, Ps_cur): directly in the outputs when full_state, else in scratch
  auto emit = [&](int64_t k, const double* m_full, const double* P_full) -> int {
    if (mo == 0) return 0;   // already written in place
    CB(gemv_rm(h->blas, false, mo, d, 1.0, Hout, d, m_full, 0.0, ms + k * mo));
    CB(gemm_rm(h->blas, false, false, mo, d, d, 1.0, Hout, d, P_full, d, 0.0, w.HPs, d));
    CB(gemm_rm(h->blas, false, true, mo, mo, d, 1.0, w.HPs, d, Hout, d, 0.0, Ps + k * (int64_t)mo * mo, mo));
    return 0;
  };
  double* m_cur = (mo == 0) ? ms + (T - 1) * d : w.ms;
  double* P_cur = (mo == 0) ? Ps + (T - 1) * dd : w.Ps0;
  CU(cudaMemcpyAsync(m_cur, mf + (T - 1) * d, (size_t)d * 8, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(P_cur, Pf + (T - 1) * dd, dd * 8, cudaMemcpyDeviceToDevice, st));
  if (int rc = emit(T - 1, m_cur, P_cur)) return rc;
  for (int64_t k = T - 2; k >= 0; --k) {
    const int64_t di = disc_index ? disc_index[k] : k;
    const double* Ak = A + di * dd;
    const double* Qk = Q + di * dd;
    const double* m_f = mf + k * d;
    const double* P_f = Pf + k * dd;
    double* m_new = (mo == 0) ? ms + k * d : w.ms;                    // vector scratch is updated in place below
    double* P_new = (mo == 0) ? Ps + k * dd : (P_cur == w.Ps0 ? w.Ps1 : w.Ps0);
    // W = Pf A^T (row-major) -- column-major this is A Pf, the right-hand side of the gain solve;
    // P_pred = A (Pf A^T) + Q
    CB(gemv_rm(h->blas, false, d, d, 1.0, Ak, d, m_f, 0.0, w.mp));
    CB(gemm_rm(h->blas, false, true, d, d, d, 1.0, P_f, d, Ak, d, 0.0, w.W, d));
    CU(cudaMemcpyAsync(w.Pp, Qk, dd * 8, cudaMemcpyDeviceToDevice, st));
    CB(gemm_rm(h->blas, false, false, d, d, d, 1.0, Ak, d, w.W, d, 1.0, w.Pp, d));
    // dP = P_next - P_pred ; dm = m_next - m_pred ; factor P_pred + jitter I
    axpby_kernel<<<blocks_for(dd), 256, 0, st>>>(w.dP, P_cur, w.Pp, -1.0, dd, 0, 0.0);
    axpby_kernel<<<blocks_for(d), 256, 0, st>>>(w.v, m_cur, w.mp, -1.0, d, 0, 0.0);
    add_diag<<<blocks_for(d), 256, 0, st>>>(w.Pp, d, jitter);
    CS(cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, d, w.Pp, d, w.lwork, w.lwork_n, w.info));
    poison_on_potrf_failure<<<1, 1, 0, st>>>(w.info, w.Pp, nullptr);
    // potrs leaves (column-major) X = (Pp + jit)^-1 A Pf, i.e. row-major X^T = G  (rts_smoother.py:58-60)
    CS(cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, d, d, w.Pp, d, w.W, d, w.info));
    // m = mf + G dm
    CU(cudaMemcpyAsync(m_new, m_f, (size_t)d * 8, cudaMemcpyDeviceToDevice, st));
    CB(gemv_rm(h->blas, false, d, d, 1.0, w.W, d, w.v, 1.0, m_new));
    // P = Pf + G dP G^T
    CB(gemm_rm(h->blas, false, false, d, d, d, 1.0, w.W, d, w.dP, d, 0.0, w.T1, d));        // G dP
    CU(cudaMemcpyAsync(P_new, P_f, dd * 8, cudaMemcpyDeviceToDevice, st));
    CB(gemm_rm(h->blas, false, true, d, d, d, 1.0, w.T1, d, w.W, d, 1.0, P_new, d));        // + (G dP) G^T
    m_cur = m_new;
    P_cur = P_new;
    if (int rc = emit(k, m_cur, P_cur)) return rc;
  }
  (void)mp_;
  CU(cudaGetLastError());
  return PHYSS_BIG_OK;
}

}  // extern "C"
