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
#include <stdint.h>

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
    // two output columns per pass: four independent accumulation chains instead of two (the DM-long dot
    // products are latency-bound with one warp per scheduler)
    int j = 0;
    for (; j + 1 < m; j += 2) {
      const double2* __restrict__ br0 = reinterpret_cast<const double2*>(B + j * LD);
      const double2* __restrict__ br1 = reinterpret_cast<const double2*>(B + (j + 1) * LD);
      double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll
      for (int l2 = 0; l2 < DM / 2; ++l2) {
        const double2 b0 = br0[l2], b1 = br1[l2];
        a00 = fma(a[2 * l2], b0.x, a00);
        a01 = fma(a[2 * l2 + 1], b0.y, a01);
        a10 = fma(a[2 * l2], b1.x, a10);
        a11 = fma(a[2 * l2 + 1], b1.y, a11);
      }
      const double base0 = Add ? Add[i * LD + j] : 0.0;
      const double base1 = Add ? Add[i * LD + j + 1] : 0.0;
      C[i * LD + j] = fma(sign, a00 + a01, base0);
      C[i * LD + j + 1] = fma(sign, a10 + a11, base1);
    }
    for (; j < m; ++j) {
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
// Compile-time block size: the lane keeps its row of A in registers, the BS entries of row j of B inside
// its diagonal block are broadcast loads (16-byte when BS is even), the row of C is written in pairs.
// All DM columns are produced (rows >= m of B are zero padding).
template <int G, int DM, int BS>
__device__ __forceinline__ void mm_nt_blk_bs(double* __restrict__ C, const double* __restrict__ A,
                                             const double* __restrict__ B, int n, const double* Add,
                                             double sign) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  for (int i = gl; i < n; i += G) {
    double a[DM];
    const double2* __restrict__ arow = reinterpret_cast<const double2*>(A + i * LD);
#pragma unroll
    for (int l2 = 0; l2 < DM / 2; ++l2) {
      const double2 t = arow[l2];
      a[2 * l2] = t.x;
      a[2 * l2 + 1] = t.y;
    }
    double2* __restrict__ crow = reinterpret_cast<double2*>(C + i * LD);
    const double2* __restrict__ addrow = Add ? reinterpret_cast<const double2*>(Add + i * LD) : nullptr;
#pragma unroll
    for (int j2 = 0; j2 < DM / 2; ++j2) {
      double o[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * j2 + h;
        const int l0 = (j / BS) * BS;
        double acc = 0.0;
        if (BS % 2 == 0) {
          const double2* __restrict__ bb = reinterpret_cast<const double2*>(B + j * LD + l0);
#pragma unroll
          for (int q = 0; q < BS / 2; ++q) {
            if (l0 + 2 * q + 1 < DM) {
              const double2 b = bb[q];
              acc = fma(a[l0 + 2 * q], b.x, acc);
              acc = fma(a[l0 + 2 * q + 1], b.y, acc);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < BS; ++q)
            if (l0 + q < DM) acc = fma(a[l0 + q], B[j * LD + l0 + q], acc);
        }
        o[h] = acc;
      }
      double2 base = addrow ? addrow[j2] : make_double2(0.0, 0.0);
      base.x = fma(sign, o[0], base.x);
      base.y = fma(sign, o[1], base.y);
      crow[j2] = base;
    }
  }
}

template <int G, int DM>
__device__ __forceinline__ void mm_nt_blk(double* __restrict__ C, const double* __restrict__ A,
                                          const double* __restrict__ B, int n, int m, int bs,
                                          const double* Add, double sign) {
  switch (bs) {
    case 1: mm_nt_blk_bs<G, DM, 1>(C, A, B, n, Add, sign); return;
    case 2: mm_nt_blk_bs<G, DM, 2>(C, A, B, n, Add, sign); return;
    case 3: mm_nt_blk_bs<G, DM, 3>(C, A, B, n, Add, sign); return;
    case 4: mm_nt_blk_bs<G, DM, 4>(C, A, B, n, Add, sign); return;
    default: break;
  }
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

// ---- FP64 tensor-core products (DMMA.8x8x4, PTX mma.sync.aligned.m8n8k4.f64) for full-size DM x DM tiles.
// Fragment layout (g = lane >> 2, t = lane & 3):
//   A (8 x 4, row)  a      = A[g][t]            B (4 x 8, col)  b = B[t][g]
//   C (8 x 8)       c0, c1 = C[g][2t], C[g][2t + 1]
// Measured on this GPU (tools/microbench/fp64_pipes.cu, profiles/fp64_pipes_r02.jsonl): DMMA.8x8x4 issues every
// 16 cycles per scheduler with 26 cycles dependent latency and runs on the SAME pipe as DFMA (no overlap, the
// same 37 TFLOP/s peak) -- it buys instruction issue slots and operand traffic (1 instruction and 2 operand
// registers per 256 multiply-adds), not arithmetic throughput.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// a[mt][kt] = M[8 mt + g][4 kt + t]: the A fragments of a DM x DM row-major slab M -- also the B fragments of M^T.
// All 32 lanes of the warp work on ONE slab (for DM < 32 the caller loops over the 32 / DM series of the warp).
template <int DM>
__device__ __forceinline__ void dmma_load_a(const double* __restrict__ M, double (&a)[DM / 8][DM / 4]) {
  constexpr int LD = Dim<DM>::LD;
  const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
#pragma unroll
  for (int mt = 0; mt < DM / 8; ++mt) {
#pragma unroll
    for (int kt = 0; kt < DM / 4; ++kt) a[mt][kt] = M[(8 * mt + g) * LD + 4 * kt + t];
  }
}

// C = Am . B   (DM x DM each), Am given as A fragments in registers, B and C row-major slabs (C must not alias B)
template <int DM>
__device__ __forceinline__ void dmma_mm_nn(double* __restrict__ C, const double (&a)[DM / 8][DM / 4],
                                           const double* __restrict__ B) {
  constexpr int LD = Dim<DM>::LD;
  constexpr int MT = DM / 8, KT = DM / 4;
  const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
  double acc[MT][MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < MT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    double b[MT];
#pragma unroll
    for (int nt = 0; nt < MT; ++nt) b[nt] = B[(4 * kt + t) * LD + 8 * nt + g];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < MT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt][kt], b[nt]);
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < MT; ++nt)
      *reinterpret_cast<double2*>(C + (8 * mt + g) * LD + 8 * nt + 2 * t) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
}

// C = Add + W . Gm^T   (DM x DM each): W, Add, C row-major slabs, Gm^T given by the A fragments of Gm
// (B[k][n] = Gm[n][k] is exactly a[nt][kt]).  C must not alias W; it may alias Add (own elements only).
template <int DM>
__device__ __forceinline__ void dmma_mm_nt(double* C, const double* __restrict__ W, const double (&ag)[DM / 8][DM / 4],
                                           const double* Add) {
  constexpr int LD = Dim<DM>::LD;
  constexpr int MT = DM / 8, KT = DM / 4;
  const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
  double acc[MT][MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < MT; ++nt) {
      const double2 v = *reinterpret_cast<const double2*>(Add + (8 * mt + g) * LD + 8 * nt + 2 * t);
      acc[mt][nt][0] = v.x;
      acc[mt][nt][1] = v.y;
    }
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) {
    double aw[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) aw[mt] = W[(8 * mt + g) * LD + 4 * kt + t];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < MT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], aw[mt], ag[nt][kt]);
  }
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < MT; ++nt)
      *reinterpret_cast<double2*>(C + (8 * mt + g) * LD + 8 * nt + 2 * t) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
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

// In-place lower Cholesky of the n x n matrix A, rd[j] = 1 / L[j][j]; returns det(A).  Non-PD -> NaN.
// `rd` must have room for 3 * LD doubles: rd[0 .. LD) the reciprocal diagonal, then two column buffers.
//
// One row per lane (G == DM): the lane keeps its row in REGISTERS.  Step j: every lane publishes its entry of
// column j to a column buffer (one store), all lanes read the pivot and the column back as 16-byte
// broadcasts and update their whole register row -- entries right of the diagonal are computed too (no
// predicates) and never used.  The strict UPPER triangle of A is filled with L^T, so that the backward
// substitution of chol_solve_t reads rows, not columns.  DM^2/2 fused multiply-adds, DM^2/4 loads and
// DM stores per factorisation, against two loads and a store per multiply-add of the in-place form.
template <int G, int DM>
__device__ __forceinline__ double chol(double* __restrict__ A, int n, double* __restrict__ rd) {
  constexpr int LD = Dim<DM>::LD;
  const int gl = lane<G>();
  static_assert(G == DM, "one row per lane");
  double det = 1.0;
  {
    const int i = gl;
    double a[DM];
    {
      const double2* __restrict__ row = reinterpret_cast<const double2*>(A + i * LD);
#pragma unroll
      for (int l2 = 0; l2 < DM / 2; ++l2) {
        const double2 t = (i < n) ? row[l2] : make_double2(0.0, 0.0);   // rows >= n may lie outside the slot
        a[2 * l2] = t.x;
        a[2 * l2 + 1] = t.y;
      }
    }
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (j < n) {                                            // uniform
        double* __restrict__ cb = rd + LD + (j & 1) * LD;
        cb[i] = a[j];                                         // column j before scaling; entry j = pivot
        __syncwarp();
        const double s = cb[j];
        const double r = fast_rsqrt(s);
        det *= s;
        const double lij = a[j] * r;
        const double f = -lij * r;                            // L[i][j] L[c][j] = (a_ij r) (a_cj r)
        const double2* __restrict__ cb2 = reinterpret_cast<const double2*>(cb);
#pragma unroll
        for (int c2 = (j + 1) / 2; c2 < DM / 2; ++c2) {
          const double2 b = cb2[c2];
          if (2 * c2 > j) a[2 * c2] = fma(f, b.x, a[2 * c2]);
          a[2 * c2 + 1] = fma(f, b.y, a[2 * c2 + 1]);
        }
        a[j] = (i == j) ? s * r : lij;
        if (i == j) rd[j] = r;
        if (i > j) A[j * LD + i] = lij;                       // L^T into the upper triangle
      }
    }
    if (i < n) {                                              // lower triangle + diagonal of the own row
      double2* __restrict__ row = reinterpret_cast<double2*>(A + i * LD);
#pragma unroll
      for (int l2 = 0; l2 < DM / 2; ++l2) {
        if (2 * l2 + 1 <= i) row[l2] = make_double2(a[2 * l2], a[2 * l2 + 1]);
        else if (2 * l2 == i) A[i * LD + i] = a[2 * l2];
      }
    }
    __syncwarp();
    return det;
  }
}

// X <- (L L^T)^{-1} X for nrhs columns, X stored TRANSPOSED: column c of the system is ROW c of Xt
// (Xt[c][0..n)), one system per lane: the lane keeps its own row in registers and the L entries are
// broadcasts.  Entries of the row beyond n must be finite (zero padding) and come back unchanged.
// The factor carries L^T in its upper triangle (chol above): both sweeps read rows.
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
        double t0 = x[i], t1 = 0.0;
        const double2* __restrict__ lrow = reinterpret_cast<const double2*>(L + i * LD);
#pragma unroll
        for (int l2 = 0; l2 < (i + 1) / 2; ++l2) {
          const double2 b = lrow[l2];
          t0 = fma(-b.x, x[2 * l2], t0);
          if (2 * l2 + 1 < i) t1 = fma(-b.y, x[2 * l2 + 1], t1);
        }
        x[i] = (t0 + t1) * rd[i];
      }
    }
#pragma unroll
    for (int i = DM - 1; i >= 0; --i) {
      if (i < n) {
        double t0 = x[i], t1 = 0.0;
        const double2* __restrict__ urow = reinterpret_cast<const double2*>(L + i * LD);
#pragma unroll
        for (int l2 = (i + 1) / 2; l2 < DM / 2; ++l2) {
          const double2 b = urow[l2];                         // zero beyond n
          if (2 * l2 > i) t0 = fma(-b.x, x[2 * l2], t0);
          t1 = fma(-b.y, x[2 * l2 + 1], t1);
        }
        x[i] = (t0 + t1) * rd[i];
      }
    }
#pragma unroll
    for (int l2 = 0; l2 < DM / 2; ++l2) row[l2] = make_double2(x[2 * l2], x[2 * l2 + 1]);
  }
}

// ---- compact block-diagonal operands (DISC_MATERN).  A block-diagonal matrix with square blocks of bs <= 4
// keeps the in-block entries of row i at Mc[i * CB + q], q < bs (CB = 4: 32-byte rows, unused entries zero);
// l0(i) = (i / bs) * bs is the first column of the block of row i.
constexpr int CB = 4;

// y[i] = sum_q Mc[i][q] x[l0(i) + q]
template <int G, int DM>
__device__ __forceinline__ void mv_c(double* __restrict__ y, const double* __restrict__ Mc,
                                     const double* __restrict__ x, int n, int bs) {
  for (int i = lane<G>(); i < n; i += G) {
    const int l0 = (i / bs) * bs;
    double acc = 0.0;
    for (int q = 0; q < bs; ++q) acc = fma(Mc[i * CB + q], x[l0 + q], acc);
    y[i] = acc;
  }
}

// M[i][l0(i) + q] += sign * Mc[i][q]   (own rows, in place)
template <int G, int DM>
__device__ __forceinline__ void add_c(double* __restrict__ M, const double* __restrict__ Mc, int n, int bs,
                                      double sign) {
  constexpr int LD = Dim<DM>::LD;
  for (int i = lane<G>(); i < n; i += G) {
    const int l0 = (i / bs) * bs;
    for (int q = 0; q < bs; ++q) M[i * LD + l0 + q] = fma(sign, Mc[i * CB + q], M[i * LD + l0 + q]);
  }
}

// Qc = Pc - Ac Pc Ac^T block by block (kernels/kernel.py:207-209, the reference's Q_k = Pinf - A Pinf A^T), all
// three compact.  The lane of row i reads the Ac / Pc rows of its own block (written by other lanes: sync first).
template <int G, int DM>
__device__ __forceinline__ void q_c(double* __restrict__ Qc, const double* __restrict__ Ac,
                                    const double* __restrict__ Pc, int n, int bs) {
  for (int i = lane<G>(); i < n; i += G) {
    const int l0 = (i / bs) * bs;
    double t[CB];
#pragma unroll
    for (int b = 0; b < CB; ++b) t[b] = 0.0;
    for (int a = 0; a < bs; ++a) {
      const double aia = Ac[i * CB + a];
#pragma unroll
      for (int b = 0; b < CB; ++b) t[b] = fma(aia, Pc[(l0 + a) * CB + b], t[b]);   // unused entries are zero
    }
    for (int q = 0; q < bs; ++q) {
      double acc = Pc[i * CB + q];
#pragma unroll
      for (int b = 0; b < CB; ++b) acc = fma(-t[b], Ac[(l0 + q) * CB + b], acc);
      Qc[i * CB + q] = acc;
    }
  }
}

// D[i][:] -= Y[i][:] over all DM columns of the own rows (in place)
template <int G, int DM>
__device__ __forceinline__ void sub_rows_inplace(double* __restrict__ D, const double* __restrict__ Y, int n) {
  constexpr int LD = Dim<DM>::LD;
  for (int i = lane<G>(); i < n; i += G) {
    const double2* __restrict__ y = reinterpret_cast<const double2*>(Y + i * LD);
    double2* __restrict__ o = reinterpret_cast<double2*>(D + i * LD);
#pragma unroll
    for (int j2 = 0; j2 < DM / 2; ++j2) {
      const double2 a = o[j2], b = y[j2];
      o[j2] = make_double2(a.x - b.x, a.y - b.y);
    }
  }
}

// D[i][:] = X[i][:] over all DM columns of the own rows
template <int G, int DM>
__device__ __forceinline__ void copy_rows(double* __restrict__ D, const double* __restrict__ X, int n) {
  constexpr int LD = Dim<DM>::LD;
  for (int i = lane<G>(); i < n; i += G) {
    const double2* __restrict__ x = reinterpret_cast<const double2*>(X + i * LD);
    double2* __restrict__ o = reinterpret_cast<double2*>(D + i * LD);
#pragma unroll
    for (int j2 = 0; j2 < DM / 2; ++j2) o[j2] = x[j2];
  }
}

// C[i][:] = sum_q Mc[i][q] B[l0(i) + q][:]  (+ Addc on the diagonal blocks)   block-diagonal times dense.
// One row per lane in registers; lanes of one block read the same rows of B.  C must not alias B.
template <int G, int DM>
__device__ __forceinline__ void mm_cn(double* __restrict__ C, const double* __restrict__ Mc,
                                      const double* __restrict__ B, int n, int bs,
                                      const double* __restrict__ Addc) {
  constexpr int LD = Dim<DM>::LD;
  for (int i = lane<G>(); i < n; i += G) {
    const int l0 = (i / bs) * bs;
    double c[DM];
#pragma unroll
    for (int j = 0; j < DM; ++j) c[j] = 0.0;
    for (int q = 0; q < bs; ++q) {
      const double a = Mc[i * CB + q];
      const double2* __restrict__ brow = reinterpret_cast<const double2*>(B + (l0 + q) * LD);
#pragma unroll
      for (int j2 = 0; j2 < DM / 2; ++j2) {
        const double2 b = brow[j2];
        c[2 * j2] = fma(a, b.x, c[2 * j2]);
        c[2 * j2 + 1] = fma(a, b.y, c[2 * j2 + 1]);
      }
    }
    double2* __restrict__ crow = reinterpret_cast<double2*>(C + i * LD);
#pragma unroll
    for (int j2 = 0; j2 < DM / 2; ++j2) crow[j2] = make_double2(c[2 * j2], c[2 * j2 + 1]);
    if (Addc)
      for (int q = 0; q < bs; ++q) C[i * LD + l0 + q] += Addc[i * CB + q];
  }
}

// C[i][j] = sum_q A[i][l0(j) + q] Bc[j][q]  (+ Addc on the diagonal blocks)   dense times block-diagonal^T.
// Block by block along the row: the BS entries of the own row inside block b and the BS x BS entries of
// block b of Bc (broadcast) give the BS outputs of the same columns, so C may be A itself (in place) and
// nothing larger than a block is live in registers.  n must be a multiple of BS.
template <int G, int DM, int BS>
__device__ __forceinline__ void mm_nc_bs(double* C, const double* A, const double* __restrict__ Bc, int n,
                                         const double* __restrict__ Addc) {
  constexpr int LD = Dim<DM>::LD;
  for (int i = lane<G>(); i < n; i += G) {
    const int bi = i / BS;
    const double* arow = A + i * LD;
    double* crow = C + i * LD;
#pragma unroll 2
    for (int b = 0; b * BS < n; ++b) {
      const int l0 = b * BS;
      double a[BS], o[BS];
      if (BS % 2 == 0) {
#pragma unroll
        for (int q = 0; q < BS / 2; ++q) {
          const double2 t = *reinterpret_cast<const double2*>(arow + l0 + 2 * q);
          a[2 * q] = t.x;
          a[2 * q + 1] = t.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < BS; ++q) a[q] = arow[l0 + q];
      }
#pragma unroll
      for (int jj = 0; jj < BS; ++jj) {
        const double* __restrict__ brow = Bc + (l0 + jj) * CB;
        double acc = 0.0;
        if (BS % 2 == 0) {
#pragma unroll
          for (int q = 0; q < BS / 2; ++q) {
            const double2 t = *reinterpret_cast<const double2*>(brow + 2 * q);
            acc = fma(a[2 * q], t.x, acc);
            acc = fma(a[2 * q + 1], t.y, acc);
          }
        } else {
#pragma unroll
          for (int q = 0; q < BS; ++q) acc = fma(a[q], brow[q], acc);
        }
        o[jj] = acc;
      }
      if (Addc && b == bi) {
#pragma unroll
        for (int jj = 0; jj < BS; ++jj) o[jj] += Addc[i * CB + jj];
      }
      if (BS % 2 == 0) {
#pragma unroll
        for (int q = 0; q < BS / 2; ++q)
          *reinterpret_cast<double2*>(crow + l0 + 2 * q) = make_double2(o[2 * q], o[2 * q + 1]);
      } else {
#pragma unroll
        for (int q = 0; q < BS; ++q) crow[l0 + q] = o[q];
      }
    }
  }
}
template <int G, int DM>
__device__ __forceinline__ void mm_nc(double* C, const double* A, const double* __restrict__ Bc, int n, int bs,
                                      const double* __restrict__ Addc) {
  switch (bs) {
    case 1: mm_nc_bs<G, DM, 1>(C, A, Bc, n, Addc); return;
    case 2: mm_nc_bs<G, DM, 2>(C, A, Bc, n, Addc); return;
    case 3: mm_nc_bs<G, DM, 3>(C, A, Bc, n, Addc); return;
    default: mm_nc_bs<G, DM, 4>(C, A, Bc, n, Addc); return;
  }
}

// dense n x m global (row stride m) <-> padded shared slot.  Full-size square matrices (n == m == DM) at a
// 16-byte aligned global address move as double2: consecutive lanes take consecutive pieces of a row, so a
// warp instruction covers whole 128-byte lines of the dense matrix.
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
template <int G, int DM>
__device__ __forceinline__ void g2s(double* __restrict__ dst, const double* __restrict__ src, int n, int m) {
  constexpr int LD = Dim<DM>::LD;
  constexpr int PR = DM / 2;
  const int gl = lane<G>();
  if (n == DM && m == DM && aligned16(src)) {
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int idx = gl; idx < DM * PR; idx += G)
      *reinterpret_cast<double2*>(dst + (idx / PR) * LD + 2 * (idx % PR)) = s2[idx];
    return;
  }
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    dst[i * LD + j] = src[idx];
  }
}
template <int G, int DM>
__device__ __forceinline__ void g2s_async(double* __restrict__ dst, const double* __restrict__ src, int n, int m) {
  constexpr int LD = Dim<DM>::LD;
  constexpr int PR = DM / 2;
  const int gl = lane<G>();
  if (n == DM && m == DM && aligned16(src)) {
#pragma unroll
    for (int idx = gl; idx < DM * PR; idx += G) cp_async16(dst + (idx / PR) * LD + 2 * (idx % PR), src + 2 * idx);
    return;
  }
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    grp::cp_async8(dst + i * LD + j, src + idx);
  }
}
template <int G, int DM>
__device__ __forceinline__ void s2g(double* __restrict__ dst, const double* __restrict__ src, int n, int m) {
  constexpr int LD = Dim<DM>::LD;
  constexpr int PR = DM / 2;
  const int gl = lane<G>();
  if (n == DM && m == DM && aligned16(dst)) {
    double2* __restrict__ d2 = reinterpret_cast<double2*>(dst);
#pragma unroll
    for (int idx = gl; idx < DM * PR; idx += G)
      d2[idx] = *reinterpret_cast<const double2*>(src + (idx / PR) * LD + 2 * (idx % PR));
    return;
  }
  for (int idx = gl; idx < n * m; idx += G) {
    const int i = idx / m, j = idx - i * m;
    dst[idx] = src[i * LD + j];
  }
}

// D[i][:] = X[i][:] - Y[i][:] over all DM columns of the own rows i < n (16-byte pieces)
template <int G, int DM>
__device__ __forceinline__ void sub_rows(double* __restrict__ D, const double* __restrict__ X,
                                         const double* __restrict__ Y, int n) {
  constexpr int LD = Dim<DM>::LD;
  for (int i = lane<G>(); i < n; i += G) {
    const double2* __restrict__ x = reinterpret_cast<const double2*>(X + i * LD);
    const double2* __restrict__ y = reinterpret_cast<const double2*>(Y + i * LD);
    double2* __restrict__ o = reinterpret_cast<double2*>(D + i * LD);
#pragma unroll
    for (int j2 = 0; j2 < DM / 2; ++j2) {
      const double2 a = x[j2], b = y[j2];
      o[j2] = make_double2(a.x - b.x, a.y - b.y);
    }
  }
}

}  // namespace rt
}  // namespace physs
