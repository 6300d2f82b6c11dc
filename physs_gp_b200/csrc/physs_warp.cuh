// physs_warp.cuh -- cooperative small dense fp64 algebra on shared-memory matrices.
//
// A "group" of G lanes (G = 8, 16 or 32, a power of two dividing the warp) owns one independent
// problem (one series / one scan element / one site block).  Matrices live in shared memory,
// row-major with an ODD leading dimension (ld = n | 1) so that column walks by consecutive lanes hit
// distinct banks.  All groups of a warp execute the same control flow (sizes are uniform across the
// launch; missing data is handled arithmetically), so phases are separated by __syncwarp().
//
// Used by the general-d sequential filter/smoother (physs_grp.cu), the parallel-in-time scan
// (physs_pscan.cu) and the CVI site kernels (physs_cvi.cu).
#pragma once
#include <cuda_runtime.h>

#include "physs_core.cuh"

namespace physs {
namespace grp {

__device__ __forceinline__ int odd_ld(int n) { return n | 1; }

template <int G>
struct Lanes {
  static __device__ __forceinline__ int gl() { return threadIdx.x & (G - 1); }
};

// C[n x m] = (ADD ? Add : 0) + sign * opA(A)[n x k] * opB(B)[k x m]
//   TA: A is stored [k x n] (use A^T);  TB: B is stored [m x k] (use B^T).
//   bsA > 0: opA(A) is block-diagonal with square blocks of size bsA (only k in the block of row i
//            contributes);  bsB > 0: opB(B) is block-diagonal (only k in the block of column j).
template <int G, bool TA, bool TB>
__device__ __forceinline__ void mm(double* C, int ldc, const double* __restrict__ A, int lda,
                                   const double* __restrict__ B, int ldb, int n, int k, int m,
                                   const double* Add, int ldadd, double sign,
                                   int bsA = 0, int bsB = 0) {
  const int gl = Lanes<G>::gl();
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    int k0 = 0, k1 = k;
    if (bsA > 0) { k0 = (i / bsA) * bsA; k1 = k0 + bsA; }
    if (bsB > 0) { k0 = (j / bsB) * bsB; k1 = k0 + bsB; }
    double acc = 0.0;
    for (int l = k0; l < k1; ++l) {
      const double a = TA ? A[l * lda + i] : A[i * lda + l];
      const double b = TB ? B[j * ldb + l] : B[l * ldb + j];
      acc = fma(a, b, acc);
    }
    C[i * ldc + j] = (Add ? Add[i * ldadd + j] : 0.0) + sign * acc;
  }
}

// y[n] = (add ? add : 0) + sign * opA(A)[n x k] x[k]
template <int G, bool TA>
__device__ __forceinline__ void mv(double* __restrict__ y, const double* __restrict__ A, int lda,
                                   const double* __restrict__ x, int n, int k,
                                   const double* __restrict__ add, double sign, int bsA = 0) {
  const int gl = Lanes<G>::gl();
  for (int i = gl; i < n; i += G) {
    int k0 = 0, k1 = k;
    if (bsA > 0) { k0 = (i / bsA) * bsA; k1 = k0 + bsA; }
    double acc = 0.0;
    for (int l = k0; l < k1; ++l) acc = fma(TA ? A[l * lda + i] : A[i * lda + l], x[l], acc);
    y[i] = (add ? add[i] : 0.0) + sign * acc;
  }
}

// In-place lower Cholesky of the n x n matrix A (only the lower triangle is read and written).
// Left-looking by columns: lane i computes entry (i, j) of column j.  Non-PD -> NaN.
// rd[j] receives 1 / L[j][j].  Returns (to every lane) the product of the squared diagonal = det(A).
template <int G>
__device__ __forceinline__ double chol(double* __restrict__ A, int ld, int n, double* __restrict__ rd) {
  const int gl = Lanes<G>::gl();
  double det = 1.0;
  for (int j = 0; j < n; ++j) {
    // diagonal
    double s = A[j * ld + j];
    for (int l = 0; l < j; ++l) s = fma(-A[j * ld + l], A[j * ld + l], s);
    const double r = fast_rsqrt(s);
    const double ljj = s * r;
    det *= s;
    __syncwarp();
    for (int i = j + gl; i < n; i += G) {
      if (i == j) {
        A[j * ld + j] = ljj;
        rd[j] = r;
      } else {
        double t = A[i * ld + j];
        for (int l = 0; l < j; ++l) t = fma(-A[i * ld + l], A[j * ld + l], t);
        A[i * ld + j] = t * r;
      }
    }
    __syncwarp();
  }
  return det;
}

// X[n x nrhs] <- (L L^T)^{-1} X, one right-hand-side column per lane (strided).
template <int G>
__device__ __forceinline__ void chol_solve(const double* __restrict__ L, int ld, int n,
                                           const double* __restrict__ rd,
                                           double* __restrict__ X, int ldx, int nrhs) {
  const int gl = Lanes<G>::gl();
  for (int c = gl; c < nrhs; c += G) {
    for (int i = 0; i < n; ++i) {
      double t = X[i * ldx + c];
      for (int l = 0; l < i; ++l) t = fma(-L[i * ld + l], X[l * ldx + c], t);
      X[i * ldx + c] = t * rd[i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double t = X[i * ldx + c];
      for (int l = i + 1; l < n; ++l) t = fma(-L[l * ld + i], X[l * ldx + c], t);
      X[i * ldx + c] = t * rd[i];
    }
  }
}

// General solve A X = B by Gaussian elimination with partial pivoting (LAPACK getrf/getrs
// semantics, which is what jsp.linalg.solve(assume_a='gen') lowers to).  A [n x n] is destroyed,
// X [n x nrhs] holds B on entry and the solution on exit.  Row operations are applied to A and X
// together, one column of [A | X] per lane.
template <int G>
__device__ __forceinline__ void lu_solve(double* __restrict__ A, int lda, int n,
                                         double* __restrict__ X, int ldx, int nrhs) {
  const int gl = Lanes<G>::gl();
  for (int j = 0; j < n; ++j) {
    // pivot search (every lane scans the column: n is small)
    int piv = j;
    double best = fabs(A[j * lda + j]);
    for (int i = j + 1; i < n; ++i) {
      const double v = fabs(A[i * lda + j]);
      if (v > best) { best = v; piv = i; }
    }
    __syncwarp();
    // swap rows j <-> piv and eliminate, column-parallel over [A(:, j+1:) | X]
    const int ncols = (n - j) + nrhs;
    const double pivval = A[piv * lda + j];
    const double rp = 1.0 / pivval;
    // multipliers depend on column j of A after the swap
    for (int c = gl; c < ncols; c += G) {
      double* col;
      int ldc;
      if (c < n - j) { col = A + (j + c); ldc = lda; } else { col = X + (c - (n - j)); ldc = ldx; }
      if (c == 0) continue;  // column j itself is handled below (needed unchanged as multipliers)
      const double top = col[piv * ldc];
      if (piv != j) { col[piv * ldc] = col[j * ldc]; col[j * ldc] = top; }
      for (int i = j + 1; i < n; ++i) {
        const double aij = (i == piv) ? A[j * lda + j] : A[i * lda + j];  // column j after the swap
        col[i * ldc] = fma(-aij * rp, top, col[i * ldc]);
      }
    }
    __syncwarp();
    // finalise column j: swap, then it is no longer read (zero below the diagonal conceptually)
    if (gl == 0) {
      if (piv != j) { A[piv * lda + j] = A[j * lda + j]; }
      A[j * lda + j] = pivval;
    }
    __syncwarp();
  }
  // back substitution, one rhs column per lane
  for (int c = gl; c < nrhs; c += G) {
    for (int i = n - 1; i >= 0; --i) {
      double t = X[i * ldx + c];
      for (int l = i + 1; l < n; ++l) t = fma(-A[i * lda + l], X[l * ldx + c], t);
      X[i * ldx + c] = t / A[i * lda + i];
    }
  }
}

// ---- asynchronous global -> shared staging (LDGSTS, 8 bytes per element: the padded shared layout
// rules out one bulk copy).  Issue, commit, and wait one step later so the copy overlaps the algebra.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int G>
__device__ __forceinline__ void g2s_async(double* __restrict__ dst, int ld,
                                          const double* __restrict__ src, int n, int m) {
  const int gl = Lanes<G>::gl();
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    cp_async8(dst + i * ld + j, src + idx);
  }
}

// copy n x m global (dense, row stride m) -> shared (ld)
template <int G>
__device__ __forceinline__ void g2s(double* __restrict__ dst, int ld, const double* __restrict__ src,
                                    int n, int m) {
  const int gl = Lanes<G>::gl();
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    dst[i * ld + j] = src[idx];
  }
}

template <int G>
__device__ __forceinline__ void s2g(double* __restrict__ dst, const double* __restrict__ src, int ld,
                                    int n, int m) {
  const int gl = Lanes<G>::gl();
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    dst[idx] = src[i * ld + j];
  }
}

// closed-form Matern transition blocks written into the diagonal blocks of A
template <int G>
__device__ __forceinline__ void matern_A(double* __restrict__ A, int ld, int d, int s, int nblk,
                                         const double* __restrict__ lam, double dt) {
  // only the diagonal blocks are written; off-block entries are never read (block-aware products)
  const int gl = Lanes<G>::gl();
  (void)d;
  for (int b = gl; b < nblk; b += G) {
    double* blk = A + (b * s) * ld + b * s;
    if (s == 1) {
      double a[1][1];
      MaternExpm<1>::eval(lam[b], dt, a);
      blk[0] = a[0][0];
    } else if (s == 2) {
      double a[2][2];
      block2_expm(lam[b], dt, a);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) blk[i * ld + j] = a[i][j];
    } else if (s == 3) {
      double a[3][3];
      MaternExpm<3>::eval(lam[b], dt, a);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) blk[i * ld + j] = a[i][j];
    } else {
      double a[4][4];
      MaternExpm<4>::eval(lam[b], dt, a);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) blk[i * ld + j] = a[i][j];
    }
  }
}

}  // namespace grp
}  // namespace physs
