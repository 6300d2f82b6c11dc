// physs_rt.cuh -- register-tiled cooperative fp64 algebra for lane groups, compile-time padded dimension.
//
// A group of G lanes owns one independent problem whose d x d matrices (d <= DM) live in shared memory,
// row-major with the fixed leading dimension LD = DM + 2 (rows 16-byte aligned), zero-padded to DM x DM.
// Lane gl owns rows gl, gl + G, ... of every product.  A row of the output is accumulated in REGISTERS
// (fully unrolled over the DM columns); the right operand is streamed row by row from shared memory with
// 16-byte loads whose address is the same for every lane of the group (a broadcast: one wavefront serves
// all groups of the warp when their slabs are staggered by 4 banks, see rt_slab).  Per fused multiply-add
// that is 1/2 .. 1/4 shared-memory load and no index arithmetic, against 2 loads + an integer division per
// element in the runtime-sized helpers of physs_warp.cuh (the instruction count per step drops ~5x).
//
// Zero padding is an invariant: rows / columns >= d of every matrix are zero (products of zero-padded
// operands stay zero-padded), so the unrolled loops may always run over all DM columns.
#pragma once
#include <cuda_runtime.h>

#include "physs_core.cuh"
#include "physs_warp.cuh"

namespace physs {
namespace rt {

template <int DM>
struct Dim {
  static constexpr int LD = DM + 2;
  static constexpr int MAT = DM * LD;            // doubles per matrix slot
  static constexpr int RB = (DM <= 16) ? 2 : 1;  // rows accumulated together per lane
};

// slab size per group: a multiple of 2 doubles, == 2 (mod 16) so that consecutive groups of a warp start
// 4 banks apart and their 16-byte broadcast loads fall into disjoint bank quads
__host__ __device__ inline int rt_slab(int doubles) {
  int t = (doubles + 1) & ~1;
  while ((t & 15) != 2) t += 2;
  return t;
}

template <int G>
__device__ __forceinline__ int lane() { return threadIdx.x & (G - 1); }

// zero a DM x LD slot
template <int G, int DM>
__device__ __forceinline__ void zero_mat(double* M) {
  for (int idx = lane<G>(); idx < Dim<DM>::MAT; idx += G) M[idx] = 0.0;
}

// C[i][:] = (Add ? Add[i][:] : 0) + sign * sum_{l in [k0, k1)} opA(A)[i][l] * B[l][:]     for rows i < n
//   TA: opA(A)[i][l] = A[l][i].   bsA > 0: opA(A) is block diagonal (square blocks of bsA): l in block of i.
// All DM columns are produced (B zero-padded => C zero-padded).  C may alias Add, must not alias A or B.
template <int G, int DM, bool TA>
__device__ __forceinline__ void mm_nn(double* __restrict__ C, const double* __restrict__ A,
                                      const double* __restrict__ B, int n, int k, const double* Add,
                                      double sign, int bsA = 0) {
  constexpr int LD = Dim<DM>::LD;
  constexpr int RB = (Dim<DM>::RB * G <= DM) ? Dim<DM>::RB : 1;   // never more row slots than rows
  const int gl = lane<G>();
  // lane gl owns the RB consecutive rows gl * RB .. gl * RB + RB - 1 of every tile of G * RB rows
  // (consecutive rows share a diagonal block, so the block range below is tight)
  for (int t0 = 0; t0 < n; t0 += G * RB) {
    const int ibase = t0 + gl * RB;
    if (ibase >= n) continue;
    int ir[RB];
    double c[RB][DM];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      ir[r] = (ibase + r < n) ? ibase + r : ibase;      // ragged last tile: recompute row `ibase`, skip the store
#pragma unroll
      for (int j = 0; j < DM; ++j) c[r][j] = 0.0;
    }
    int k0 = 0, k1 = k;
    if (bsA > 0) {
      // opA(A) is stored as a full matrix with zeros off its diagonal blocks, so the range only has to
      // COVER the blocks of the tile's rows
      k0 = (ir[0] / bsA) * bsA;
      k1 = (ir[RB - 1] / bsA) * bsA + bsA;
      if (k1 > k) k1 = k;
    }
    for (int l = k0; l < k1; ++l) {
      double a[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) a[r] = TA ? A[l * LD + ir[r]] : A[ir[r] * LD + l];
      const double2* __restrict__ brow = reinterpret_cast<const double2*>(B + l * LD);
#pragma unroll
      for (int j2 = 0; j2 < DM / 2; ++j2) {
        const double2 b = brow[j2];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          c[r][2 * j2] = fma(a[r], b.x, c[r][2 * j2]);
          c[r][2 * j2 + 1] = fma(a[r], b.y, c[r][2 * j2 + 1]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int i = ibase + r;
      if (i < n) {
        double2* __restrict__ crow = reinterpret_cast<double2*>(C + i * LD);
        const double2* __restrict__ arow = Add ? reinterpret_cast<const double2*>(Add + i * LD) : nullptr;
#pragma unroll
        for (int j2 = 0; j2 < DM / 2; ++j2) {
          double2 o = arow ? arow[j2] : make_double2(0.0, 0.0);
          o.x = fma(sign, c[r][2 * j2], o.x);
          o.y = fma(sign, c[r][2 * j2 + 1], o.y);
          crow[j2] = o;
        }
      }
    }
  }
}

// C[i][j] = (Add ? Add[i][j] : 0) + sign * sum_l A[i][l] * B[j][l]   for i < n, j < m   (dense A B^T)
// The lane keeps its row of A in registers; rows of B are broadcast.  Columns >= m of C are left untouched.
template <int G, int DM, bool TA = false>
__device__ __forceinline__ void mm_nt(double* __restrict__ C, const double* __restrict__ A,
                                      const double* __restrict__ B, int n, int m, const double* Add,
                                      double sign) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int i = gl; i < n; i += G) {
    double a[DM];
    if (TA) {                                             // row i of A^T = column i of A
#pragma unroll
      for (int l = 0; l < DM; ++l) a[l] = A[l * LD + i];
    } else {
      const double2* __restrict__ arow = reinterpret_cast<const double2*>(A + i * LD);
#pragma unroll
      for (int l2 = 0; l2 < DM / 2; ++l2) {
        const double2 t = arow[l2];
        a[2 * l2] = t.x;
        a[2 * l2 + 1] = t.y;
      }
    }
    for (int j = 0; j < m; ++j) {
      const double2* __restrict__ brow = reinterpret_cast<const double2*>(B + j * LD);
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
      for (int l2 = 0; l2 < DM / 2; ++l2) {
        const double2 b = brow[l2];
        acc0 = fma(a[2 * l2], b.x, acc0);
        acc1 = fma(a[2 * l2 + 1], b.y, acc1);
      }
      const double base = Add ? Add[i * LD + j] : 0.0;
      C[i * LD + j] = fma(sign, acc0 + acc1, base);
    }
  }
}

// C[i][j] = Add[i][j] + sign * sum_{l in block(j)} A[i][l] * B[j][l]   (A opB with block-diagonal opB = B^T)
template <int G, int DM>
__device__ __forceinline__ void mm_nt_blk(double* __restrict__ C, const double* __restrict__ A,
                                          const double* __restrict__ B, int n, int m, int bs,
                                          const double* Add, double sign) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int i = gl; i < n; i += G) {
    for (int j = 0; j < m; ++j) {
      const int l0 = (j / bs) * bs;
      double acc = 0.0;
      for (int l = l0; l < l0 + bs; ++l) acc = fma(A[i * LD + l], B[j * LD + l], acc);
      C[i * LD + j] = fma(sign, acc, Add ? Add[i * LD + j] : 0.0);
    }
  }
}

// y[i] = (add ? add[i] : 0) + sign * sum_l opA(A)[i][l] x[l]
template <int G, int DM, bool TA>
__device__ __forceinline__ void mv(double* __restrict__ y, const double* __restrict__ A,
                                   const double* __restrict__ x, int n, int k, const double* __restrict__ add,
                                   double sign, int bsA = 0) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int i = gl; i < n; i += G) {
    int k0 = 0, k1 = k;
    if (bsA > 0) { k0 = (i / bsA) * bsA; k1 = k0 + bsA; }
    double acc = 0.0;
    for (int l = k0; l < k1; ++l) acc = fma(TA ? A[l * LD + i] : A[i * LD + l], x[l], acc);
    y[i] = fma(sign, acc, add ? add[i] : 0.0);
  }
}

// In-place lower Cholesky of the n x n matrix A (lower triangle), rd[j] = 1 / L[j][j]; returns det(A).
// Right-looking by columns: after column j is scaled, lane i updates its trailing row i with the
// broadcast column.  Non-PD -> NaN.
template <int G, int DM>
__device__ __forceinline__ double chol(double* __restrict__ A, int n, double* __restrict__ rd) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  double det = 1.0;
  for (int j = 0; j < n; ++j) {
    const double s = A[j * LD + j];
    const double r = fast_rsqrt(s);
    det *= s;
    __syncwarp();
    // column j: L[i][j] = A[i][j] * r  (i > j), diagonal = s * r
    for (int i = j + gl; i < n; i += G) {
      if (i == j) { A[j * LD + j] = s * r; rd[j] = r; }
      else A[i * LD + j] *= r;
    }
    __syncwarp();
    // trailing update: A[i][c] -= L[i][j] L[c][j]  for j < c <= i
    for (int i = j + 1 + gl; i < n; i += G) {
      const double lij = A[i * LD + j];
      for (int c = j + 1; c <= i; ++c) A[i * LD + c] = fma(-lij, A[c * LD + j], A[i * LD + c]);
    }
    __syncwarp();
  }
  return det;
}

// X <- (L L^T)^{-1} X for nrhs columns, X stored TRANSPOSED: column c of the system is ROW c of Xt
// (Xt[c][0..n)), one system per lane: the lane keeps its own row in registers and the L entries are
// broadcasts.  Entries of the row beyond n are passed through unchanged.
template <int G, int DM>
__device__ __forceinline__ void chol_solve_t(const double* __restrict__ L, int n, const double* __restrict__ rd,
                                             double* __restrict__ Xt, int nrhs) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int c = gl; c < nrhs; c += G) {
    double x[DM];
    double2* __restrict__ row = reinterpret_cast<double2*>(Xt + c * LD);
#pragma unroll
    for (int l2 = 0; l2 < DM / 2; ++l2) {
      const double2 t = row[l2];
      x[2 * l2] = t.x;
      x[2 * l2 + 1] = t.y;
    }
    // forward: x[i] = (x[i] - sum_{l<i} L[i][l] x[l]) * rd[i]; the loops are unrolled over DM (x stays in
    // registers) and every row beyond n is skipped by a uniform branch
#pragma unroll
    for (int i = 0; i < DM; ++i) {
      if (i < n) {
        double t = x[i];
        const double2* __restrict__ lrow = reinterpret_cast<const double2*>(L + i * LD);
#pragma unroll
        for (int l2 = 0; l2 < (i + 1) / 2; ++l2) {
          const double2 b = lrow[l2];
          t = fma(-b.x, x[2 * l2], t);
          if (2 * l2 + 1 < i) t = fma(-b.y, x[2 * l2 + 1], t);
        }
        x[i] = t * rd[i];
      }
    }
#pragma unroll
    for (int i = DM - 1; i >= 0; --i) {
      if (i < n) {
        double t = x[i];
#pragma unroll
        for (int l = i + 1; l < DM; ++l) {
          if (l < n) t = fma(-L[l * LD + i], x[l], t);
        }
        x[i] = t * rd[i];
      }
    }
#pragma unroll
    for (int l2 = 0; l2 < DM / 2; ++l2) row[l2] = make_double2(x[2 * l2], x[2 * l2 + 1]);
  }
}

// dense n x m global (row stride m) <-> padded shared slot
template <int G, int DM>
__device__ __forceinline__ void g2s(double* __restrict__ dst, const double* __restrict__ src, int n, int m) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    dst[i * LD + j] = src[idx];
  }
}
template <int G, int DM>
__device__ __forceinline__ void g2s_async(double* __restrict__ dst, const double* __restrict__ src, int n, int m) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    grp::cp_async8(dst + i * LD + j, src + idx);
  }
}
template <int G, int DM>
__device__ __forceinline__ void s2g(double* __restrict__ dst, const double* __restrict__ src, int n, int m) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    dst[idx] = src[i * LD + j];
  }
}

}  // namespace rt
}  // namespace physs
