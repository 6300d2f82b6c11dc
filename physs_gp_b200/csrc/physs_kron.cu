// physs_kron.cu -- hand-written large-block Kalman filter / RTS smoother for ONE series of a separable
// spatio-temporal prior (BASELINE config 2: Matern-3/2 (time) x RBF (space), Ns = 200 -> d = 400, m = 200).
//
// Reference: kernels/kernel.py:213-265 and kernels/ss_utils.py:41-53 -- A_k = I (x) A_t, P_inf = K_s (x) P_inf_t,
// H = I (x) [1 0 ..] (state index = s * ds + j) -- driven through kalman_filter.py:144-241,439-485 and
// rts_smoother.py:48-106,162-192.  The reference builds the dense d x d Kronecker matrices every step and
// multiplies them out (2 d^3 per predict); here the structure is used:
//
//   predict           P_ = A P A^T + Q          ds x ds block transforms, O(d^2)          (elementwise phase)
//   innovation        S = M P_[::ds, ::ds] M + R a strided gather                          (same phase)
//   gain              X = W L^-T, K = X L^-1    blocked Cholesky + blocked substitutions (L L^T = S + jitter I)
//   covariance        P = P_ - K S K^T = P_ - X X^T + jitter K K^T       one rank-2m update on the tensor cores
//   lml               second Cholesky of the un-jittered, mask-to-identity S (gaussian.py:72-108)
//
// One PERSISTENT cooperative kernel runs all T filter steps: 1 CTA per SM, a two-level atomic grid barrier between
// the phases of a step -- innovation Cholesky (shared-memory resident, with look-ahead), gain solves (factor resident
// in shared memory, 8 rows per CTA), rank-2m covariance update with the predict of the NEXT step fused into its
// epilogue.  Every dense product is a tile GEMM on DMMA.8x8x4 (mma.sync.m8n8k4.f64): warps 0-3 issue the DMMAs,
// warps 4-7 stage the next 128-deep operand chunk with 16-byte cp.async.  The smoother splits into a part that is
// PARALLEL over time -- the gains G_k = Pf_k A^T (A Pf_k A^T + Q + jitter I)^-1 depend on the filter output only: one
// CTA per time step, blocked Cholesky + two blocked triangular solves out of L2 -- and the sequential recursion
// P_s,k = Pf_k + G_k (P_s,k+1 - P_pred,k) G_k^T, two distributed GEMMs per step in a second persistent cooperative
// kernel.  The same shared-memory Cholesky serves the CVI site update / surrogate ELL of D <= 208 site blocks
// (kron_cvi_site_kernel, kron_cvi_ell_sur_kernel).  Measured phase budgets: DESIGN.md section 3, "Config 2".
//
// All matrices row-major fp64; NaN on numerical failure (sqrt of a non-positive pivot), status 0.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "physs_core.cuh"
#include "physs_internal.h"

namespace physs {
namespace kron {

constexpr int NTH = 256;      // threads per CTA: 8 warps = 4 column strips x 2 k-halves
constexpr int TN = 32;        // tile width
constexpr int KC = 128;       // largest k-chunk staged per pipeline stage
constexpr int NST = 2;        // pipeline stages (measured: 2 x 128 beats 4 x 64 -- the per-chunk cost is instruction issue
                              // on 8 warps, not operand latency)
constexpr int LDA = KC + 4;   // 132: (LDA mod 16) == 4 -> the 8 x 4 fragment loads of a half-warp hit 16 distinct bank pairs
constexpr int LDB = TN + 4;   // 36: same property for the 4 x 8 fragment loads of a [K x N] tile
constexpr int TMMAX = 40;
constexpr int SM_A = NST * TMMAX * LDA;
constexpr int SM_B = NST * KC * LDB;           // KC * LDB = 2304 >= TN * LDA = 2176 (an [N x K] tile fits as well)
constexpr int SM_RED = 4 * 5 * 2 * 32;         // split-k partial sums; reused as the 32 x 33 transposition tile
constexpr int SM_DG = 32 * 33 + 32;
constexpr int SM_DOUBLES = SM_A + SM_B + SM_RED + SM_DG;
constexpr size_t SM_BYTES = SM_DOUBLES * sizeof(double);

__device__ __forceinline__ void cp_async16(double* dst, const double* src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4 g + t holds a = A[g][t], b = B[t][g], c = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// accumulate the nanoseconds since *tick into prof[slot] (thread 0 of CTA 0 only; measurement aid)
__device__ __forceinline__ void ptick(double* prof, int slot, unsigned long long& tick) {
  if (prof && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long now = gtimer();
    prof[slot] += (double)(now - tick);
    tick = now;
  }
}

// ---------------------------------------------------------------------------------- grid-wide barrier
// Two-level arrive counter (groups of 16 CTAs, then one top counter) with monotonic 64-bit counts: atomics on one
// address serialise at ~27 cycles each, so 148 arrivals on a single counter cost ~2 us; 16 + 10 cost a quarter.
// The kernels are launched cooperatively (all CTAs resident).  bar: [0] top counter, [16 (g + 1)] group counters
// (one 128-byte line each), zeroed before the launch.  __threadfence() on both sides of the arrive / wait is the
// release / acquire (it also invalidates this SM's L1, so plain loads after the barrier see the other CTAs' data).
constexpr int BAR_GROUP = 16;
struct GridBarrier {
  unsigned long long* bar;
  unsigned long long round;
  __device__ __forceinline__ void init(unsigned long long* b) { bar = b; round = 0; }
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      ++round;
      const unsigned ncta = gridDim.x, grp = blockIdx.x / BAR_GROUP;
      const unsigned ngrp = (ncta + BAR_GROUP - 1) / BAR_GROUP;
      const unsigned gsize = min((unsigned)BAR_GROUP, ncta - grp * BAR_GROUP);
      __threadfence();
      const unsigned long long old = atomicAdd(bar + 16 * (grp + 1), 1ULL);
      if (old + 1 == round * gsize) atomicAdd(bar, 1ULL);
      const unsigned long long target = round * ngrp;
      unsigned long long seen;
      do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(bar) : "memory");
      } while (seen < target);
      __threadfence();
    }
    __syncthreads();
  }
};

// ------------------------------------------------------------------------------------------ tile GEMM
// C[M x N] = Add + sum_p sg[p] * A_p[M x K_p] * op(B_p),  op(B) = B ([K x N] row-major, nt = 0) or B^T (B is
// [N x K] row-major, nt = 1).  Output tiles TM x 32 (TM = 8 MT) are dealt round-robin to `ncta` CTAs
// (cta = 0, ncta = 1: the whole product on this CTA).  lower = 1 (needs M == N, MT == 4): only tiles on or
// below the diagonal are computed and mirrored into the upper triangle (through a shared-memory transpose, so
// both copies are written with full rows).  Optional second outputs: C2 = C - Sub2 (same shape) and the
// projection Cp[i / ds][j / ds] = C[i][j] for i % ds == j % ds == 0.  With two operand pairs sg[1] must be != 0.
struct Gemm {
  const double* A[2]; int lda[2];
  const double* B[2]; int ldb[2];
  int K[2]; double sg[2]; int npair;
  int nt;
  int M, N;
  const double* Add; int ldadd;
  double* C; int ldc;
  int lower;
  double* C2; const double* Sub2; int ld2;
  double* Cp; int ldp; int ds;
  double* prof;                 // measurement aid: slots 24.. (thread 0 of CTA 0)
  const struct PredictFuse* fuse;   // filter only: predict of the NEXT step applied to every finished tile
  const double* add_scale;      // optional device scalar multiplying Add (null: 1)
};

// The predict phase of step k + 1 fused into the update epilogue of step k (lower mode, 32 % ds == 0): a finished
// 32 x 32 tile of P_k holds whole ds x ds blocks, so  P_ = A P A^T + Q,  the innovation covariances and the masked
// P_ H^T of the next step follow tile by tile -- one phase and one grid barrier less per step.  The outputs go to the
// OTHER half of the double-buffered (Pp, W): this step's tiles are still being read by other CTAs.
struct PredictFuse {
  int Ns, ds, d;
  const double* At; const double* Qt;      // ds x ds of step k + 1
  const double* Ks; const double* y; const double* R; double jitter;
  double* Pp; double* W; double* SjA; double* SmA;
};

__device__ __forceinline__ bool vec_ok(const double* p, int64_t ld) {
  return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld & 1) == 0);
}
// Warp roles inside gemm_tiles: warps 0-3 (one per scheduler) issue the DMMAs, each on one 8-column strip of the
// tile with MT independent accumulator pairs; warps 4-7 only stage the NEXT k-chunk with cp.async while the
// current one is consumed.  (All eight warps doing both put ~900 address / predicate instructions per chunk in
// front of 80 DMMAs on every warp: measured 3.3 us per 40 x 32 x 128 chunk against 1.35 us of DMMA issue time.)
constexpr int NLOAD = 128;     // loader threads (tid >= NTH - NLOAD)
// [ROWS x kpad] block of a row-major matrix -> shared memory (leading dim LDA), zero outside the valid extent
// vr x vc.  lt = loader thread index: 64 threads x 16 bytes cover one row of up to 128 doubles, 2 rows per pass.
template <int ROWS>
__device__ __forceinline__ void load_rows(int lt, double* dst, const double* src, int64_t ld, int kpad, int vr, int vc,
                                          bool vec) {
  const int c = (lt & 63) << 1;
  if (c >= kpad) return;
  int r = lt >> 6;
  const double* s = src + (int64_t)r * ld + c;
  double* d = dst + r * LDA + c;
  if (vec && vr >= ROWS && vc >= kpad) {           // interior tile: no predicates
#pragma unroll
    for (int i = 0; i < ROWS / 2; ++i) cp_async16(d + i * 2 * LDA, s + (int64_t)i * 2 * ld);
    return;
  }
#pragma unroll
  for (int i = 0; i < ROWS / 2; ++i, r += 2, s += 2 * ld, d += 2 * LDA) {
    if (vec && r < vr && c + 1 < vc) {
      cp_async16(d, s);
    } else {
      const double x0 = (r < vr && c < vc) ? __ldcg(s) : 0.0;
      const double x1 = (r < vr && c + 1 < vc) ? __ldcg(s + 1) : 0.0;
      *reinterpret_cast<double2*>(d) = make_double2(x0, x1);
    }
  }
}
// [kpad x 32] block ([K x N] operand) -> shared memory (leading dim LDB): 16 threads per row, 8 rows per pass
__device__ __forceinline__ void load_kn(int lt, double* dst, const double* src, int64_t ld, int kpad, int vr, int vc,
                                        bool vec) {
  const int c = (lt & 15) << 1;
  if (vec && vr >= kpad && vc >= TN) {             // interior tile: no predicates
    const double* s = src + (int64_t)(lt >> 4) * ld + c;
    double* d = dst + (lt >> 4) * LDB + c;
    for (int r = lt >> 4; r < kpad; r += 8, s += 8 * ld, d += 8 * LDB) cp_async16(d, s);
    return;
  }
  for (int r = lt >> 4; r < kpad; r += 8) {
    const double* s = src + (int64_t)r * ld + c;
    double* d = dst + r * LDB + c;
    if (vec && r < vr && c + 1 < vc) {
      cp_async16(d, s);
    } else {
      const double x0 = (r < vr && c < vc) ? __ldcg(s) : 0.0;
      const double x1 = (r < vr && c + 1 < vc) ? __ldcg(s + 1) : 0.0;
      *reinterpret_cast<double2*>(d) = make_double2(x0, x1);
    }
  }
}

constexpr int DSMAX = 4;

__device__ __noinline__ void fused_predict(const PredictFuse& f, const double* Tt, int r0, int c0, bool mirror) {
  const int ds = f.ds, Ns = f.Ns, d = f.d, m = f.Ns;
  const int nbt = 32 / ds;
  double at[DSMAX][DSMAX], qt[DSMAX][DSMAX];
#pragma unroll
  for (int a = 0; a < DSMAX; ++a)
#pragma unroll
    for (int b = 0; b < DSMAX; ++b) {
      at[a][b] = (a < ds && b < ds) ? f.At[a * ds + b] : 0.0;
      qt[a][b] = (a < ds && b < ds) ? f.Qt[a * ds + b] : 0.0;
    }
  for (int bp = threadIdx.x; bp < nbt * nbt; bp += NTH) {
    const int bi = bp / nbt, bj = bp - bi * nbt;
    const int I = r0 / ds + bi, J = c0 / ds + bj;
    if (I >= Ns || J >= Ns) continue;
    double blk[DSMAX][DSMAX], tmp[DSMAX][DSMAX];
#pragma unroll
    for (int a = 0; a < DSMAX; ++a)
#pragma unroll
      for (int b = 0; b < DSMAX; ++b) blk[a][b] = (a < ds && b < ds) ? Tt[(bi * ds + a) * 33 + bj * ds + b] : 0.0;
#pragma unroll
    for (int a = 0; a < DSMAX; ++a)
#pragma unroll
      for (int b = 0; b < DSMAX; ++b) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < DSMAX; ++c) s = fma(at[a][c], blk[c][b], s);
        tmp[a][b] = s;
      }
    const double ks = f.Ks[(int64_t)I * Ns + J];
    const double yi = f.y[I], yj = f.y[J];
    const bool oi = !(yi != yi), oj = !(yj != yj);
#pragma unroll
    for (int a = 0; a < DSMAX; ++a)
#pragma unroll
      for (int b = 0; b < DSMAX; ++b) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < DSMAX; ++c) s = fma(tmp[a][c], at[b][c], s);
        blk[a][b] = s + ks * qt[a][b];
        if (a < ds && b < ds) {
          f.Pp[(int64_t)(I * ds + a) * d + J * ds + b] = blk[a][b];
          if (mirror) f.Pp[(int64_t)(J * ds + b) * d + I * ds + a] = blk[a][b];
        }
      }
    const double p00 = (oi && oj) ? blk[0][0] : 0.0;
    {
      const double sv = p00 + f.R[(int64_t)I * Ns + J];
      f.SjA[(int64_t)I * m + J] = sv + (I == J ? f.jitter : 0.0);
      f.SmA[(int64_t)I * m + J] = (oi && oj) ? sv : (I == J ? 1.0 : 0.0);
    }
    if (mirror) {                                     // I != J in a tile strictly below the diagonal
      const double sv = p00 + f.R[(int64_t)J * Ns + I];
      f.SjA[(int64_t)J * m + I] = sv;
      f.SmA[(int64_t)J * m + I] = (oi && oj) ? sv : 0.0;
    }
#pragma unroll
    for (int a = 0; a < DSMAX; ++a)
      if (a < ds) {
        f.W[(int64_t)(I * ds + a) * m + J] = oj ? blk[a][0] : 0.0;
        if (mirror) f.W[(int64_t)(J * ds + a) * m + I] = oi ? blk[0][a] : 0.0;      // P_ symmetric
      }
  }
}

template <int MT>
__device__ __noinline__ void gemm_tiles(const Gemm& g, int cta, int ncta, double* sm) {
  constexpr int TM = 8 * MT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const bool loader = tid >= NTH - NLOAD;
  const int lt = tid - (NTH - NLOAD);
  const int tm = (g.M + TM - 1) / TM, tn = (g.N + TN - 1) / TN;
  const int ntiles = g.lower ? tm * (tm + 1) / 2 : tm * tn;
  // k-chunks: as few as fit KC, all of (nearly) the same length, a multiple of 4
  const int nch0 = (g.K[0] + KC - 1) / KC;
  const int kch0 = ((g.K[0] + nch0 - 1) / nch0 + 3) & ~3;
  const int nch1 = g.npair > 1 ? (g.K[1] + KC - 1) / KC : 0;
  const int kch1 = nch1 ? ((g.K[1] + nch1 - 1) / nch1 + 3) & ~3 : 0;
  const int nch = nch0 + nch1;
  const int my_tiles = ntiles > cta ? (ntiles - cta + ncta - 1) / ncta : 0;
  const int nitems = my_tiles * nch;
  if (nitems == 0) return;
  double* As = sm;
  double* Bs = sm + SM_A;
  double* Tt = sm + SM_A + SM_B;
  const bool va0 = vec_ok(g.A[0], g.lda[0]), vb0 = vec_ok(g.B[0], g.ldb[0]);
  const bool va1 = g.npair > 1 && vec_ok(g.A[1], g.lda[1]), vb1 = g.npair > 1 && vec_ok(g.B[1], g.ldb[1]);

  auto tile_origin = [&](int tile_no, int& r0, int& c0) {
    const int t = cta + tile_no * ncta;
    if (g.lower) {
      int r = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((r + 1) * (r + 2) / 2 <= t) ++r;
      while (r * (r + 1) / 2 > t) --r;
      r0 = r * TM;
      c0 = (t - r * (r + 1) / 2) * TN;
    } else {
      r0 = (t / tn) * TM;
      c0 = (t % tn) * TN;
    }
  };
  auto chunk_of = [&](int ch, int& pr, int& k0, int& kpad) {
    pr = ch >= nch0 ? 1 : 0;
    if (pr) ch -= nch0;
    const int kch = pr ? kch1 : kch0;
    k0 = ch * kch;
    const int kc = max(0, min(kch, g.K[pr] - k0));
    kpad = (kc + 3) & ~3;
  };
  int ld_tile = 0, ld_ch = 0, ld_r0, ld_c0;      // loaders: the (tile, chunk) the next issue() stages
  tile_origin(0, ld_r0, ld_c0);
  auto issue = [&](int buf) {
    int pr, k0, kpad;
    chunk_of(ld_ch, pr, k0, kpad);
    const int K = g.K[pr];
    load_rows<TM>(lt, As + buf * TMMAX * LDA, g.A[pr] + (int64_t)ld_r0 * g.lda[pr] + k0, g.lda[pr], kpad, g.M - ld_r0,
                  K - k0, pr ? va1 : va0);
    if (g.nt)
      load_rows<TN>(lt, Bs + buf * KC * LDB, g.B[pr] + (int64_t)ld_c0 * g.ldb[pr] + k0, g.ldb[pr], kpad, g.N - ld_c0,
                    K - k0, pr ? vb1 : vb0);
    else
      load_kn(lt, Bs + buf * KC * LDB, g.B[pr] + (int64_t)k0 * g.ldb[pr] + ld_c0, g.ldb[pr], kpad, K - k0, g.N - ld_c0,
              pr ? vb1 : vb0);
    cp_async_commit();
    if (++ld_ch == nch) {
      ld_ch = 0;
      ++ld_tile;
      if (ld_tile < my_tiles) tile_origin(ld_tile, ld_r0, ld_c0);
    }
  };

  double acc[MT][2];
  int r0, c0;
  tile_origin(0, r0, c0);
  if (loader) {                 // two chunks in flight from the start
    issue(0);
    if (nitems > 1) issue(1);
    else cp_async_commit();
    cp_async_wait<1>();
  }
  double addv[MT][2];
  int ch = 0, tile_no = 0;
  unsigned long long tk = gtimer();
  for (int it = 0; it < nitems; ++it) {
    int pr, k0, kpad;
    chunk_of(ch, pr, k0, kpad);
    ptick(g.prof, 24, tk);
    __syncthreads();            // chunk `it` has landed (the loaders waited for it); chunk it - 1 has been consumed
    ptick(g.prof, it == 0 ? 25 : 26, tk);
    if (loader) {
      if (it >= 1 && it + 1 < nitems) issue((it + 1) & 1);
      cp_async_wait<0>();
    } else {
      if (ch == nch - 1 && g.Add) {                  // the addend of the epilogue: loads in flight during the k-loop
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int row = r0 + 8 * mt + gq, col = c0 + 8 * warp + 2 * tq;
          addv[mt][0] = (row < g.M && col < g.N) ? __ldcg(g.Add + (int64_t)row * g.ldadd + col) : 0.0;
          addv[mt][1] = (row < g.M && col + 1 < g.N) ? __ldcg(g.Add + (int64_t)row * g.ldadd + col + 1) : 0.0;
        }
      }
      if (ch == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
      } else if (ch == nch0) {
        const double f = g.sg[0] / g.sg[1];        // second operand pair: the final scale is sg[1]
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { acc[mt][0] *= f; acc[mt][1] *= f; }
      }
      const double* as = As + (it & 1) * TMMAX * LDA + gq * LDA + tq;
      const double* bs = Bs + (it & 1) * KC * LDB + (g.nt ? (8 * warp + gq) * LDA + tq : tq * LDB + 8 * warp + gq);
      const int bstep = g.nt ? 4 : 4 * LDB;
      const int ksteps = kpad >> 2;
      // software-pipelined: the fragments of k-step kt + 1 are loaded before the DMMAs of k-step kt are issued
      // (a warp issues in order; with the loads behind the DMMAs the FP64 pipe idles for a shared-memory
      // round trip every step -- measured 29 instead of 16 cycles per DMMA)
      double a0[MT], a1[MT], b0, b1;
      if (ksteps > 0) {
        b0 = bs[0];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a0[mt] = as[8 * mt * LDA];
      }
      int kt = 0;
      for (; kt + 2 <= ksteps; kt += 2) {
        b1 = bs[(kt + 1) * bstep];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a1[mt] = as[8 * mt * LDA + 4 * kt + 4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][0], acc[mt][1], a0[mt], b0);
        if (kt + 2 < ksteps) {
          b0 = bs[(kt + 2) * bstep];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) a0[mt] = as[8 * mt * LDA + 4 * kt + 8];
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][0], acc[mt][1], a1[mt], b1);
      }
      if (kt < ksteps) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][0], acc[mt][1], a0[mt], b0);
      }
    }
    ptick(g.prof, 27, tk);
    if (ch == nch - 1) {
      const double sgl = g.sg[g.npair - 1];
      const bool mirror = g.lower && r0 > c0;
      if (!loader) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int row = r0 + 8 * mt + gq;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c0 + 8 * warp + 2 * tq + e;
            double v = sgl * acc[mt][e];
            if (row < g.M && col < g.N) {
              if (g.Add) v = g.add_scale ? fma(__ldg(g.add_scale), addv[mt][e], v) : v + addv[mt][e];
              g.C[(int64_t)row * g.ldc + col] = v;
              if (g.C2) g.C2[(int64_t)row * g.ld2 + col] = v - __ldcg(g.Sub2 + (int64_t)row * g.ld2 + col);
              if (g.Cp && row % g.ds == 0 && col % g.ds == 0) g.Cp[(int64_t)(row / g.ds) * g.ldp + col / g.ds] = v;
            }
            acc[mt][e] = v;
          }
        }
      }
      if (mirror || g.fuse) {                         // MT == 4 here: stage the finished 32 x 32 tile in shared memory
        __syncthreads();                              // (Tt may still be read by the previous tile's readers)
        if (!loader) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            Tt[(8 * mt + gq) * 33 + 8 * warp + 2 * tq] = acc[mt][0];
            Tt[(8 * mt + gq) * 33 + 8 * warp + 2 * tq + 1] = acc[mt][1];
          }
        }
        __syncthreads();
        if (g.fuse) fused_predict(*g.fuse, Tt, r0, c0, mirror);
        for (int idx = tid; mirror && idx < 1024; idx += NTH) {
          const int cc = idx >> 5, rr = idx & 31;      // consecutive lanes: consecutive columns of the mirrored row
          const int row = c0 + cc, col = r0 + rr;
          if (row < g.N && col < g.M) {
            const double v = Tt[rr * 33 + cc];
            g.C[(int64_t)row * g.ldc + col] = v;
            if (g.C2) g.C2[(int64_t)row * g.ld2 + col] = v - __ldcg(g.Sub2 + (int64_t)row * g.ld2 + col);
            if (g.Cp && row % g.ds == 0 && col % g.ds == 0) g.Cp[(int64_t)(row / g.ds) * g.ldp + col / g.ds] = v;
          }
        }
      }
      ch = 0;
      if (++tile_no < my_tiles) tile_origin(tile_no, r0, c0);
    } else {
      ++ch;
    }
  }
  ptick(g.prof, 28, tk);
  __syncthreads();              // the staging buffers are free again
}

// ------------------------------------------------------------------------- 32 x 32 diagonal blocks
// Factor the w x w block at Lm in place (lower Cholesky factor; strictly upper part of the block zeroed).
// Leaves the factor padded with the identity in Dg (ld 33) and the reciprocal diagonal in rd[32].
// Warp 0, lane i = row i held in registers; column j is scaled by rsqrt(pivot) and broadcast with shuffles, so
// the chain through the 32 pivots is shuffle + rsqrt + multiply-add with no shared-memory round trip.
__device__ __noinline__ void diag_factor(double* Lm, int64_t ld, int w, double* Dg, double* rd) {
  if (threadIdx.x < 32) {
    const int i = threadIdx.x;
    double r[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) r[c] = (i < w && c <= i) ? __ldcg(Lm + i * ld + c) : (c == i ? 1.0 : 0.0);
    double rdi = 1.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double dj = __shfl_sync(0xffffffffu, r[j], j);
      const double rs = fast_rsqrt(dj);            // dj <= 0 -> NaN (the reference's Cholesky of a non-PD matrix)
      const double lij = r[j] * rs;                // lane j: sqrt(dj); lanes > j: L[i][j]; lanes < j: 0
      r[j] = lij;
      if (i == j) rdi = rs;
#pragma unroll
      for (int c = j + 1; c < 32; ++c) {
        const double lcj = __shfl_sync(0xffffffffu, lij, c);
        if (c <= i) r[c] = fma(-lij, lcj, r[c]);
      }
    }
    rd[i] = rdi;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      Dg[i * 33 + c] = r[c];
      if (i < w && c < w) Lm[i * ld + c] = (c <= i) ? r[c] : 0.0;
    }
  }
  __syncthreads();
}
// Load an already factored diagonal block (identity padding) and its reciprocal diagonal.
__device__ __noinline__ void diag_load(const double* Lm, int64_t ld, int w, double* Dg, double* rd) {
  for (int idx = threadIdx.x; idx < 1024; idx += NTH) {
    const int r = idx >> 5, c = idx & 31;
    Dg[r * 33 + c] = (r < w && c <= r) ? __ldcg(Lm + r * ld + c) : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  if (threadIdx.x < 32) rd[threadIdx.x] = 1.0 / Dg[threadIdx.x * 33 + threadIdx.x];
  __syncthreads();
}
// One thread per row: x L^T = t (forward) or x L = t (backward) against the padded 32 x 32 factor in Dg, in
// place on nr rows of X (columns [0, w)); optionally duplicated into X2.
template <bool FWD>
__device__ __noinline__ void subst_rows(double* X, int64_t ldx, int nr, int w, const double* Dg, const double* rd, double* X2,
                           int64_t ldx2) {
  for (int r = threadIdx.x; r < nr; r += NTH) {
    double t[32];
    double* xr = X + (int64_t)r * ldx;
#pragma unroll
    for (int c = 0; c < 32; ++c) t[c] = (c < w) ? __ldcg(xr + c) : 0.0;
    if (FWD) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const double xk = t[k] * rd[k];
        t[k] = xk;
#pragma unroll
        for (int c = k + 1; c < 32; ++c) t[c] = fma(-xk, Dg[c * 33 + k], t[c]);
      }
    } else {
#pragma unroll
      for (int k = 31; k >= 0; --k) {
        const double xk = t[k] * rd[k];
        t[k] = xk;
#pragma unroll
        for (int c = 0; c < k; ++c) t[c] = fma(-xk, Dg[k * 33 + c], t[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (c < w) {
        xr[c] = t[c];
        if (X2) X2[(int64_t)r * ldx2 + c] = t[c];
      }
    }
  }
}

// ------------------------------------------------------------ CTA-local blocked factorisation / solves
// Left-looking blocked Cholesky of the n x n matrix at Lm (lower triangle referenced and overwritten), carried
// through `nrows` >= n rows: extra rows r >= n come out as r L^-T (an appended right-hand side v^T gives
// (L^-1 v)^T).  One CTA; everything stays in L2 / global memory, tiles pass through shared memory.
__device__ __noinline__ void chol_blocked(double* Lm, int64_t ld, int n, int nrows, double* sm, double* prof = nullptr) {
  double* Dg = sm + SM_A + SM_B + SM_RED;
  double* rd = Dg + 32 * 33;
  unsigned long long tk = gtimer();
  for (int r0 = 0; r0 < n; r0 += 32) {
    const int w = min(32, n - r0);
    if (r0 > 0) {
      Gemm g = {};
      g.npair = 1; g.nt = 1;
      g.A[0] = Lm + (int64_t)r0 * ld; g.lda[0] = (int)ld;
      g.B[0] = Lm + (int64_t)r0 * ld; g.ldb[0] = (int)ld;
      g.K[0] = r0; g.sg[0] = -1.0;
      g.M = nrows - r0; g.N = w;
      g.Add = Lm + (int64_t)r0 * ld + r0; g.ldadd = (int)ld;
      g.C = Lm + (int64_t)r0 * ld + r0; g.ldc = (int)ld;
      gemm_tiles<5>(g, 0, 1, sm);
      __syncthreads();
    }
    ptick(prof, 16, tk);
    diag_factor(Lm + (int64_t)r0 * ld + r0, ld, w, Dg, rd);
    __syncthreads();
    ptick(prof, 17, tk);
    const int nr = nrows - (r0 + w);
    if (nr > 0) subst_rows<true>(Lm + (int64_t)(r0 + w) * ld + r0, ld, nr, w, Dg, rd, nullptr, 0);
    __syncthreads();
    ptick(prof, 18, tk);
  }
}
// X L^T = W in place on R rows of X (n columns), L the factor left by chol_blocked.
template <int MT>
__device__ __noinline__ void trsm_fwd(double* X, int64_t ldx, int R, const double* Lm, int64_t ldl, int n, double* sm,
                         double* prof = nullptr) {
  double* Dg = sm + SM_A + SM_B + SM_RED;
  double* rd = Dg + 32 * 33;
  unsigned long long tk = gtimer();
  for (int r0 = 0; r0 < n; r0 += 32) {
    const int w = min(32, n - r0);
    if (r0 > 0) {
      Gemm g = {};
      g.npair = 1; g.nt = 1;
      g.A[0] = X; g.lda[0] = (int)ldx;
      g.B[0] = Lm + (int64_t)r0 * ldl; g.ldb[0] = (int)ldl;
      g.K[0] = r0; g.sg[0] = -1.0;
      g.M = R; g.N = w;
      g.Add = X + r0; g.ldadd = (int)ldx;
      g.C = X + r0; g.ldc = (int)ldx;
      gemm_tiles<MT>(g, 0, 1, sm);
      __syncthreads();
    }
    ptick(prof, 20, tk);
    diag_load(Lm + (int64_t)r0 * ldl + r0, ldl, w, Dg, rd);
    ptick(prof, 21, tk);
    subst_rows<true>(X + r0, ldx, R, w, Dg, rd, nullptr, 0);
    __syncthreads();
    ptick(prof, 22, tk);
  }
}
// G L = Y in place on R rows of X; the result is also written to X2 when given.
template <int MT>
__device__ __noinline__ void trsm_bwd(double* X, int64_t ldx, int R, const double* Lm, int64_t ldl, int n, double* X2, int64_t ldx2,
                         double* sm) {
  double* Dg = sm + SM_A + SM_B + SM_RED;
  double* rd = Dg + 32 * 33;
  const int nb = (n + 31) / 32;
  for (int j = nb - 1; j >= 0; --j) {
    const int r0 = 32 * j, w = min(32, n - r0), r1 = r0 + w;
    if (r1 < n) {
      Gemm g = {};
      g.npair = 1; g.nt = 0;
      g.A[0] = X + r1; g.lda[0] = (int)ldx;
      g.B[0] = Lm + (int64_t)r1 * ldl + r0; g.ldb[0] = (int)ldl;
      g.K[0] = n - r1; g.sg[0] = -1.0;
      g.M = R; g.N = w;
      g.Add = X + r0; g.ldadd = (int)ldx;
      g.C = X + r0; g.ldc = (int)ldx;
      gemm_tiles<MT>(g, 0, 1, sm);
      __syncthreads();
    }
    diag_load(Lm + (int64_t)r0 * ldl + r0, ldl, w, Dg, rd);
    subst_rows<false>(X + r0, ldx, R, w, Dg, rd, X2 ? X2 + r0 : nullptr, ldx2);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- shared-memory resident variants (n <= 208)
// The per-step m x m innovation Cholesky and the gain solves sit on the critical path of the filter; run out of
// L2 they are a chain of ~100 dependent round trips.  For m <= 208 the whole lower triangle fits in shared memory
// (8-row block rows, block row I holds 8 (I + 1) columns with leading dimension 8 (I + 1) + 4: the 8 x 4 fragment
// loads of a half-warp then hit 16 distinct bank pairs in every block row):
constexpr int TRI_MAXBR = 26;
__device__ __forceinline__ int tri_off(int I) { return 32 * I * (I + 2); }
__device__ __forceinline__ int tri_ld(int I) { return 8 * (I + 1) + 4; }
__device__ __forceinline__ int tri_idx(int r, int c) { const int I = r >> 3; return tri_off(I) + (r & 7) * tri_ld(I) + c; }
constexpr int TRI_DOUBLES = 32 * TRI_MAXBR * (TRI_MAXBR + 2);      // 23296
constexpr int TRI_N = 8 * TRI_MAXBR;                               // 208
// layout of the shared-memory arena in these phases (doubles)
constexpr int TS_L = 0;                          // triangle
constexpr int TS_DG = TS_L + TRI_DOUBLES;        // 32 x 33 padded diagonal block + 32 reciprocal pivots
constexpr int TS_V = TS_DG + SM_DG;              // appended right-hand side (TRI_N)
constexpr int TS_X = TS_V + TRI_N + 32;               // 8 rows of X (ld TRI_N + 4), filter phase 3
constexpr int TS_T = TS_X + 8 * (TRI_N + 4);     // 8 x 36 tile + 2 x 4 x 64 split-k partials
constexpr int TS_DOUBLES = TS_T + 8 * 36 + 512;
static_assert(TS_DOUBLES * 8 <= 227 * 1024, "shared-memory arena of the resident factorisation too large");
constexpr size_t SM_BYTES_FILTER = (TS_DOUBLES > SM_DOUBLES ? TS_DOUBLES : SM_DOUBLES) * sizeof(double);

// global (row-major, ld) lower triangle -> shared triangle; rows >= n are identity padding
__device__ void tri_load_issue(double* Ls, const double* Sg, int64_t ld, int n, int nbr) {
  // one warp per block row, lane = (row in the block row, one of four interleaved 16-byte column units): no
  // divisions, every lane busy whatever the row length
  const bool vec = vec_ok(Sg, ld);
  const int lane = threadIdx.x & 31, rl = lane >> 2, sub = lane & 3;
  for (int I = threadIdx.x >> 5; I < nbr; I += NTH / 32) {
    const int r = 8 * I + rl;
    double* d = Ls + tri_off(I) + rl * tri_ld(I);
    const double* s = Sg + (int64_t)r * ld;
    for (int c = 2 * sub; c < 8 * (I + 1); c += 8) {
      if (r < n && vec && c + 1 < n) {
        cp_async16(d + c, s + c);
      } else {
        const double x0 = (r < n && c < n) ? __ldcg(s + c) : (r == c ? 1.0 : 0.0);
        const double x1 = (r < n && c + 1 < n) ? __ldcg(s + c + 1) : (r == c + 1 ? 1.0 : 0.0);
        *reinterpret_cast<double2*>(d + c) = make_double2(x0, x1);
      }
    }
  }
  cp_async_commit();
}
__device__ void tri_load(double* Ls, const double* Sg, int64_t ld, int n, int nbr) {
  tri_load_issue(Ls, Sg, ld, n, nbr);
  cp_async_wait<0>();
  __syncthreads();
}

// warp 0: factor the 32 x 32 (w32 valid) diagonal block at column c0 of the shared triangle in place; padded
// copy in Dg, reciprocal pivots in rd.  Lane i holds row i in registers; column j goes through a 32-entry
// shared buffer (one store, then broadcast 16-byte loads) -- per pivot 1 STS + <= 16 LDS instead of 62 shuffles,
// which a single warp issues at ~5 cycles each.  col: 2 x 32 doubles (double-buffered, 16-byte aligned).
__device__ __noinline__ void tri_diag_factor(double* Ls, int c0, int w32, double* Dg, double* rd, double* col) {
  const int i = threadIdx.x;
  double r[32];
  const int base = (i < w32) ? tri_idx(c0 + i, c0) : 0;
#pragma unroll
  for (int c = 0; c < 32; ++c) r[c] = (i < w32 && c <= i) ? Ls[base + c] : (c == i ? 1.0 : 0.0);
  double rdi = 1.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    double* cb = col + (j & 1) * 32;
    const double aij = r[j];
    cb[i] = aij;                                   // column j below the diagonal (0 for i < j)
    __syncwarp();
    const double dj = cb[j];
    const double rs = fast_rsqrt(dj);              // dj <= 0 -> NaN (the reference's Cholesky of a non-PD matrix)
    const double t = aij * (rs * rs);              // a_ij / d_j
    r[j] = aij * rs;                               // l_ij (lane j: sqrt(d_j))
    if (i == j) rdi = rs;
#pragma unroll
    for (int c2 = (j + 1) >> 1; c2 < 16; ++c2) {
      // unconditional: entries above the diagonal (c > i) turn into garbage that nothing reads -- the column
      // buffer is only read at rows >= j -- and that is zeroed on write-back; a predicate costs two selects each
      const double2 q = reinterpret_cast<const double2*>(cb)[c2];
      if (2 * c2 > j) r[2 * c2] = fma(-t, q.x, r[2 * c2]);
      r[2 * c2 + 1] = fma(-t, q.y, r[2 * c2 + 1]);
    }
  }
  rd[i] = rdi;
  const int wid = 8 * (((c0 + i) >> 3) + 1) - c0;      // stored columns of this row inside the block
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    Dg[i * 33 + c] = (c <= i) ? r[c] : 0.0;
    if (i < w32 && c < wid) Ls[base + c] = (c <= i) ? r[c] : 0.0;
  }
}
// one thread per row: x L_JJ^T = t on the 32 columns at c0 of rows [row0, row0 + nr) of the shared triangle, and
// (thread nr) of the appended vector v
__device__ __noinline__ void tri_panel_solve(double* Ls, double* v, int c0, int row0, int nr, const double* Dg,
                                             const double* rd) {
  const int t_ = threadIdx.x;
  if (t_ > nr) return;
  double* x = (t_ < nr) ? Ls + tri_idx(row0 + t_, c0) : v + c0;
  double t[32];
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const double2 q = *reinterpret_cast<const double2*>(x + c);
    t[c] = q.x; t[c + 1] = q.y;
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const double xk = t[k] * rd[k];
    t[k] = xk;
#pragma unroll
    for (int c = k + 1; c < 32; ++c) t[c] = fma(-xk, Dg[c * 33 + k], t[c]);
  }
#pragma unroll
  for (int c = 0; c < 32; c += 2) *reinterpret_cast<double2*>(x + c) = make_double2(t[c], t[c + 1]);
}

// Cholesky of the n x n matrix Sg (+ appended row n = right-hand side v) entirely in shared memory, n <= 208.
// On return (after a __syncthreads) the shared triangle holds L, v holds L^-1 v.  Lg != nullptr: L (lower
// triangle) and the solved row are written back to global memory in the layout of Sg.
__device__ __noinline__ void chol_smem(const double* Sg, int64_t ld, int n, double* Lg, double* sm, double* prof) {
  double* Ls = sm + TS_L;
  double* Dg = sm + TS_DG;
  double* rd = Dg + 32 * 33;
  double* v = sm + TS_V;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int nbr = (n + 7) >> 3, npad = 8 * nbr;
  unsigned long long tk = gtimer();
  tri_load_issue(Ls, Sg, ld, n, nbr);
  for (int c = tid; c < npad + 32; c += NTH) v[c] = c < n ? __ldcg(Sg + (int64_t)n * ld + c) : 0.0;
  cp_async_wait<0>();
  __syncthreads();
  ptick(prof, 16, tk);
  // four 8 x 8 tiles (I, K0 .. K0 + 3) of the trailing update C[I][K] -= L[I][c0..] L[K][c0..]^T: independent
  // accumulator chains (DMMA latency 26 cycles > issue interval 16)
  auto quad = [&](int I, int K0, int c0, int ks) {
    double a[8];
    const double* ai = Ls + tri_idx(8 * I + gq, c0 + tq);
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) a[kt] = kt < ks ? -ai[4 * kt] : 0.0;
    const double* bk[4];
    double* cp[4];
    double2 cc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int Kq = min(K0 + q, I);
      bk[q] = Ls + tri_idx(8 * Kq + gq, c0 + tq);
      cp[q] = Ls + tri_idx(8 * I + gq, 8 * Kq + 2 * tq);
      cc[q] = *reinterpret_cast<double2*>(cp[q]);
    }
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      if (kt < ks) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(cc[q].x, cc[q].y, a[kt], bk[q][4 * kt]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (K0 + q <= I) *reinterpret_cast<double2*>(cp[q]) = cc[q];
  };
  // Right-looking, 32 columns per step, with look-ahead: once the NEXT block column is up to date, warp 0 factors
  // its diagonal block (a chain of 32 dependent pivots, ~6 us) while warps 1-7 finish the trailing update.
  if (warp == 0) tri_diag_factor(Ls, 0, min(32, npad), Dg, rd, sm + TS_T);
  __syncthreads();
  ptick(prof, 17, tk);
  for (int c0 = 0; c0 < npad; c0 += 32) {
    const int w32 = min(32, npad - c0);
    const int row0 = c0 + w32, nr = npad - row0;
    tri_panel_solve(Ls, v, c0, row0, nr, Dg, rd);
    __syncthreads();
    ptick(prof, 18, tk);
    if (nr > 0) {
      const int I0 = row0 >> 3, ks = w32 >> 2;
      for (int I = I0 + warp; I < nbr; I += NTH / 32) quad(I, I0, c0, ks);      // next block column first
      for (int c = row0 + tid; c < npad; c += NTH) {                            // appended vector
        const double* lc = Ls + tri_idx(c, c0);
        double s = v[c];
        for (int k = 0; k < w32; ++k) s = fma(-v[c0 + k], lc[k], s);
        v[c] = s;
      }
      __syncthreads();
      ptick(prof, 19, tk);
      if (warp == 0) {
        tri_diag_factor(Ls, row0, min(32, npad - row0), Dg, rd, sm + TS_T);
      } else {
        int cnt = 0;
        for (int I = I0 + 4; I < nbr; ++I)
          for (int K0 = I0 + 4; K0 <= I; K0 += 4)
            if (cnt++ % (NTH / 32 - 1) == warp - 1) quad(I, K0, c0, ks);
      }
      __syncthreads();
      ptick(prof, 17, tk);
    }
  }
  if (Lg) {
    for (int r = warp; r < n; r += NTH / 32) {
      const double* s = Ls + tri_idx(r, 0);
      for (int c = lane; c <= r; c += 32) Lg[(int64_t)r * ld + c] = s[c];
    }
    for (int c = tid; c < n; c += NTH) Lg[(int64_t)n * ld + c] = v[c];
  }
}

// Filter phase 3 on one CTA: R <= 8 rows of X = W L^-T (in place in global Wr, ld = n) and, when Kr != nullptr,
// of K = X L^-1, with the whole factor L resident in shared memory.  The 32 x 32 diagonal blocks of L are
// inverted once (one warp per block, in place), so every block step is DMMA work: an update with the columns
// already solved and a multiply by the inverse block.  Also returns the mean update  out_m[i] = mp[i] + X[i,:] . u.
// prepared: the factor (with inverted diagonal blocks) is still resident from a previous call on this CTA;
// ident_row0 >= 0: the right-hand side is rows ident_row0 .. + R of the identity (rows of (L L^T)^-1 come out in Kr).
__device__ __noinline__ void trsm_smem(const double* Lg, int64_t ld, int n, double* Wr, int R, double* Kr, const double* u,
                          const double* mp, double* out_m, double* sm, double* prof, bool prepared = false,
                          int ident_row0 = -1) {
  unsigned long long tk = gtimer();
  double* Ls = sm + TS_L;
  double* Xs = sm + TS_X;
  double* Ts = sm + TS_T;
  double* Part = Ts + 8 * 36;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int nt = warp & 3, half = warp >> 2;
  const int nbr = (n + 7) >> 3, npad = 8 * nbr, nb32 = (npad + 31) >> 5;
  constexpr int LDX = TRI_N + 4;
  if (!prepared) tri_load_issue(Ls, Lg, ld, n, nbr);
  for (int idx = tid; idx < 8 * npad; idx += NTH) {
    const int r = idx / npad, c = idx - r * npad;
    if (ident_row0 >= 0) Xs[r * LDX + c] = (r < R && c == ident_row0 + r) ? 1.0 : 0.0;
    else Xs[r * LDX + c] = (r < R && c < n) ? __ldcg(Wr + (int64_t)r * ld + c) : 0.0;
  }
  cp_async_wait<0>();
  __syncthreads();
  ptick(prof, 20, tk);
  // invert the diagonal blocks in place: warp J takes block J; lane c owns column c of the inverse
  for (int J = warp; !prepared && J < nb32; J += NTH / 32) {
    const int c0 = 32 * J, w32 = min(32, npad - c0);
    const int c = lane;
    double x[32];
    const double dinv = (c < w32) ? 1.0 / Ls[tri_idx(c0 + c, c0 + c)] : 1.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const double rdi = __shfl_sync(0xffffffffu, dinv, i);
      double s = (i == c) ? 1.0 : 0.0;
      if (i < w32) {
        const double* li = Ls + tri_idx(c0 + i, c0);
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;                     // four partial sums: the dot product is a chain
#pragma unroll
        for (int k = 0; k < i; ++k) {                             // x[k] == 0 for k < c
          if ((k & 3) == 0) s = fma(-li[k], x[k], s);
          else if ((k & 3) == 1) s1 = fma(-li[k], x[k], s1);
          else if ((k & 3) == 2) s2 = fma(-li[k], x[k], s2);
          else s3 = fma(-li[k], x[k], s3);
        }
        s = (s + s1) + (s2 + s3);
      }
      x[i] = (i >= c) ? s * rdi : 0.0;
    }
    __syncwarp();
    if (c < w32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < w32) {
          const int wid = 8 * (((c0 + i) >> 3) + 1) - c0;
          if (c < wid) Ls[tri_idx(c0 + i, c0 + c)] = x[i];       // zero above the diagonal inside the stored width
        }
      }
    }
  }
  __syncthreads();
  ptick(prof, 21, tk);
  // forward: X L^T = W, block columns left to right
  for (int J = 0; J < nb32; ++J) {
    const int c0 = 32 * J, w32 = min(32, npad - c0);
    const bool act = 8 * nt < w32;
    double acc0 = 0.0, acc1 = 0.0;
    if (act) {
      const double* xa = Xs + gq * LDX + tq;
      const double* lb = Ls + tri_idx(c0 + 8 * nt + gq, tq);
      double e0 = 0.0, e1 = 0.0;                       // second accumulator chain (DMMA latency 26 > issue 16)
      int kt = half;
      for (; kt + 2 < 8 * J; kt += 4) {
        dmma884(acc0, acc1, xa[4 * kt], lb[4 * kt]);
        dmma884(e0, e1, xa[4 * kt + 8], lb[4 * kt + 8]);
      }
      if (kt < 8 * J) dmma884(acc0, acc1, xa[4 * kt], lb[4 * kt]);
      acc0 += e0; acc1 += e1;
    }
    if (half == 1) { Part[(nt * 2 + 0) * 32 + lane] = acc0; Part[(nt * 2 + 1) * 32 + lane] = acc1; }
    __syncthreads();
    if (half == 0 && act) {
      const double2 w = *reinterpret_cast<const double2*>(Xs + gq * LDX + c0 + 8 * nt + 2 * tq);
      Ts[gq * 36 + 8 * nt + 2 * tq] = w.x - (acc0 + Part[(nt * 2 + 0) * 32 + lane]);
      Ts[gq * 36 + 8 * nt + 2 * tq + 1] = w.y - (acc1 + Part[(nt * 2 + 1) * 32 + lane]);
    }
    __syncthreads();
    if (half == 0 && act) {
      // X_J[:, n-tile nt] = sum_k T[:, k] inv[n][k], inv lower triangular: k < 8 (nt + 1)
      double o0 = 0.0, o1 = 0.0, p0 = 0.0, p1 = 0.0;
      const double* ta = Ts + gq * 36 + tq;
      const double* ib = Ls + tri_idx(c0 + 8 * nt + gq, c0 + tq);
      for (int kt = 0; kt < 2 * (nt + 1); kt += 2) {       // the trip count is even
        dmma884(o0, o1, ta[4 * kt], ib[4 * kt]);
        dmma884(p0, p1, ta[4 * kt + 4], ib[4 * kt + 4]);
      }
      *reinterpret_cast<double2*>(Xs + gq * LDX + c0 + 8 * nt + 2 * tq) = make_double2(o0 + p0, o1 + p1);
    }
    __syncthreads();
  }
  // X rows back to global, mean update
  for (int idx = tid; Wr && ident_row0 < 0 && idx < R * n; idx += NTH) {
    const int r = idx / n, c = idx - r * n;
    Wr[(int64_t)r * ld + c] = Xs[r * LDX + c];
  }
  if (u && warp < R) {
    double s = 0.0;
    for (int a = lane; a < n; a += 32) s = fma(Xs[warp * LDX + a], __ldcg(u + a), s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out_m[warp] = mp[warp] + s;
  }
  ptick(prof, 22, tk);
  if (!Kr) return;
  __syncthreads();
  // backward: K L = X, block columns right to left (in place in Xs)
  for (int J = nb32 - 1; J >= 0; --J) {
    const int c0 = 32 * J, w32 = min(32, npad - c0), c1 = c0 + w32;
    const bool act = 8 * nt < w32;
    double acc0 = 0.0, acc1 = 0.0;
    if (act) {
      // sum over solved columns k >= c1: K[:, k] L[k][c0 + n]
      const double* xa = Xs + gq * LDX + c1 + tq;
      const int ksteps = (npad - c1) >> 2;
      double e0 = 0.0, e1 = 0.0;
      int kt = half;
      for (; kt + 2 < ksteps; kt += 4) {
        dmma884(acc0, acc1, xa[4 * kt], Ls[tri_idx(c1 + 4 * kt + tq, c0 + 8 * nt + gq)]);
        dmma884(e0, e1, xa[4 * kt + 8], Ls[tri_idx(c1 + 4 * kt + 8 + tq, c0 + 8 * nt + gq)]);
      }
      if (kt < ksteps) dmma884(acc0, acc1, xa[4 * kt], Ls[tri_idx(c1 + 4 * kt + tq, c0 + 8 * nt + gq)]);
      acc0 += e0; acc1 += e1;
    }
    if (half == 1) { Part[(nt * 2 + 0) * 32 + lane] = acc0; Part[(nt * 2 + 1) * 32 + lane] = acc1; }
    __syncthreads();
    if (half == 0 && act) {
      const double2 w = *reinterpret_cast<const double2*>(Xs + gq * LDX + c0 + 8 * nt + 2 * tq);
      Ts[gq * 36 + 8 * nt + 2 * tq] = w.x - (acc0 + Part[(nt * 2 + 0) * 32 + lane]);
      Ts[gq * 36 + 8 * nt + 2 * tq + 1] = w.y - (acc1 + Part[(nt * 2 + 1) * 32 + lane]);
    }
    __syncthreads();
    if (half == 0 && act) {
      // K_J[:, n-tile nt] = sum_k T[:, k] inv[k][n], inv lower triangular: k >= 8 nt
      double o0 = 0.0, o1 = 0.0, p0 = 0.0, p1 = 0.0;
      const double* ta = Ts + gq * 36 + tq;
      for (int kt = 2 * nt; kt < (w32 >> 2); kt += 2) {    // w32 is a multiple of 8: the trip count is even
        dmma884(o0, o1, ta[4 * kt], Ls[tri_idx(c0 + 4 * kt + tq, c0 + 8 * nt + gq)]);
        dmma884(p0, p1, ta[4 * kt + 4], Ls[tri_idx(c0 + 4 * kt + 4 + tq, c0 + 8 * nt + gq)]);
      }
      *reinterpret_cast<double2*>(Xs + gq * LDX + c0 + 8 * nt + 2 * tq) = make_double2(o0 + p0, o1 + p1);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < R * n; idx += NTH) {
    const int r = idx / n, c = idx - r * n;
    Kr[(int64_t)r * ld + c] = Xs[r * LDX + c];
  }
  ptick(prof, 23, tk);
}

// block reduction of one value over the CTA (result valid in thread 0)
__device__ double block_sum(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < NTH / 32; ++i) s += scratch[i];
  return s;
}


// ----------------------------------------------------------------- CVI site update / surrogate ELL, large blocks
// The D = m = Ns site blocks of the config-2 CVI posterior (FullConjugateGaussian over the spatial points of one time
// step; cvi_nat_grad.py:47-87, exponential_family_transforms.py:25-95, expected_log_likelihoods.py:90-117): one CTA per
// time step, D <= 208.  Every SPD inverse is the shared-memory Cholesky above followed by the resident gain solves on
// identity rows (rows of (L L^T)^-1, eight at a time).
struct CviBigArgs {
  int64_t T; int D;
  const double* Yt; const double* Vt;          // sites [T, D], [T, D, D]
  const double* qm; const double* qS;          // posterior marginals of the site blocks (qS: surrogate ELL only)
  const double* dm; const double* dS;          // dELL/dm [T, D], dELL/dS [T, D, D] or its diagonal [T, D]
  int dS_diag;
  double beta, ngj;
  double* Yn; double* Vn;                      // updated sites
  double* ell;                                 // surrogate ELL per step [T]
  double* scratch;                             // per CTA: 2 (D + 1) D + D D + 2 D doubles
};
__host__ __device__ __forceinline__ int64_t cvi_big_scratch(int D) { return 2 * (int64_t)(D + 1) * D + (int64_t)D * D + 2 * D + 8; }

// (M + 0)^-1 for the SPD matrix in Mg ((D + 1) x D, appended row = right-hand side v): factor into Lg, rows of the
// inverse into Inv (ld D).  On return v holds L^-1 v in shared memory (TS_V) and the diagonal of L is in the triangle.
__device__ void spd_inverse_big(const double* Mg, double* Lg, double* Inv, int D, double* sm) {
  chol_smem(Mg, D, D, Lg, sm, nullptr);
  __syncthreads();
  for (int rb = 0; rb * 8 < D; ++rb) {
    const int R = min(8, D - 8 * rb);
    trsm_smem(Lg, D, D, nullptr, R, Inv + (int64_t)rb * 8 * D, nullptr, nullptr, nullptr, sm, nullptr, rb > 0, 8 * rb);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NTH, 1) kron_cvi_site_kernel(const CviBigArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, D = p.D;
  const int64_t DD = (int64_t)D * D;
  double* Mg = p.scratch + (int64_t)blockIdx.x * cvi_big_scratch(D);
  double* Lg = Mg + (int64_t)(D + 1) * D;
  double* Inv = Lg + (int64_t)(D + 1) * D;
  double* l1 = Inv + DD;
  for (int64_t t = blockIdx.x; t < p.T; t += gridDim.x) {
    const double* Yt = p.Yt + t * D;
    const double* Vt = p.Vt + t * DD;
    const double* qm = p.qm + t * D;
    const double* dm = p.dm + t * D;
    const double* dS = p.dS + t * (p.dS_diag ? (int64_t)D : DD);
    // theta -> lambda: Vinv = (V~ + ng_jitter I)^-1
    for (int idx = tid; idx < (D + 1) * D; idx += NTH) {
      const int i = idx / D, j = idx - i * D;
      Mg[idx] = i < D ? Vt[idx] + (i == j ? p.ngj : 0.0) : 0.0;
    }
    __syncthreads();
    spd_inverse_big(Mg, Lg, Inv, D, sm);
    // block update: lambda_1' = (1 - beta) Vinv Y~ + beta (dm - 2 dS m),  -2 lambda_2' = (1 - beta) Vinv - 2 beta dS
    for (int i = warp; i < D; i += NTH / 32) {
      double a = 0.0, g = 0.0;
      for (int k = lane; k < D; k += 32) {
        a = fma(__ldcg(Inv + (int64_t)i * D + k), Yt[k], a);
        if (!p.dS_diag) g = fma(dS[(int64_t)i * D + k], qm[k], g);
      }
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); g += __shfl_xor_sync(0xffffffffu, g, o); }
      if (lane == 0) {
        if (p.dS_diag) g = dS[i] * qm[i];
        l1[i] = (1.0 - p.beta) * a + p.beta * (dm[i] - 2.0 * g);
      }
    }
    for (int idx = tid; idx < (D + 1) * D; idx += NTH) {
      const int i = idx / D, j = idx - i * D;
      if (i < D) {
        const double ds = p.dS_diag ? (i == j ? dS[i] : 0.0) : dS[idx];
        Mg[idx] = (1.0 - p.beta) * __ldcg(Inv + idx) - 2.0 * p.beta * ds + (i == j ? p.ngj : 0.0);
      } else {
        Mg[idx] = 0.0;
      }
    }
    __syncthreads();
    // lambda -> theta: V~' = (-2 lambda_2' + ng_jitter I)^-1,  Y~' = V~' lambda_1'
    double* Vn = p.Vn + t * DD;
    spd_inverse_big(Mg, Lg, Vn, D, sm);
    for (int i = warp; i < D; i += NTH / 32) {
      double a = 0.0;
      for (int k = lane; k < D; k += 32) a = fma(__ldcg(Vn + (int64_t)i * D + k), __ldcg(l1 + k), a);
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) p.Yn[t * D + i] = a;
    }
    __syncthreads();
  }
}

// surrogate ELL of one step: log N(Y~ | m, V~) - 1/2 tr(V~^-1 S)   (sites carry no missing entries)
__global__ void __launch_bounds__(NTH, 1) kron_cvi_ell_sur_kernel(const CviBigArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, D = p.D;
  const int64_t DD = (int64_t)D * D;
  double* Mg = p.scratch + (int64_t)blockIdx.x * cvi_big_scratch(D);
  double* Lg = Mg + (int64_t)(D + 1) * D;
  double* Inv = Lg + (int64_t)(D + 1) * D;
  for (int64_t t = blockIdx.x; t < p.T; t += gridDim.x) {
    const double* Yt = p.Yt + t * D;
    const double* Vt = p.Vt + t * DD;
    const double* qm = p.qm + t * D;
    const double* qS = p.qS + t * DD;
    for (int idx = tid; idx < (D + 1) * D; idx += NTH) {
      const int i = idx / D, j = idx - i * D;
      Mg[idx] = i < D ? Vt[idx] : Yt[j] - qm[j];
    }
    __syncthreads();
    chol_smem(Mg, D, D, Lg, sm, nullptr);
    __syncthreads();
    double ld = 0.0, mh = 0.0, tr = 0.0;
    for (int a = tid; a < D; a += NTH) {
      ld += log(sm[TS_L + tri_idx(a, a)]);
      const double u = sm[TS_V + a];
      mh = fma(u, u, mh);
    }
    __syncthreads();
    for (int rb = 0; rb * 8 < D; ++rb) {
      const int R = min(8, D - 8 * rb);
      trsm_smem(Lg, D, D, nullptr, R, Inv + (int64_t)rb * 8 * D, nullptr, nullptr, nullptr, sm, nullptr, rb > 0, 8 * rb);
      __syncthreads();
    }
    for (int idx = tid; idx < D * D; idx += NTH) {
      const int i = idx / D, j = idx - i * D;
      tr = fma(__ldcg(Inv + idx), qS[(int64_t)j * D + i], tr);
    }
    double* scr = sm + TS_T;
    ld = block_sum(ld, scr);
    mh = block_sum(mh, scr);
    tr = block_sum(tr, scr);
    if (tid == 0) p.ell[t] = -0.5 * (D * 1.8378770664093454835606594728112 + 2.0 * ld + mh + tr);
    __syncthreads();
  }
  (void)warp; (void)lane;
}

// phase timers (ns, CTA 0): accumulated into the workspace tail when PHYSS_KRON_PROF is set -- measurement aid
#define KRON_TICK(slot)                                              \
  if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) {               \
    const unsigned long long now_ = gtimer();                        \
    p.prof[slot] += (double)(now_ - tick_);                          \
    tick_ = now_;                                                    \
  }

// ---------------------------------------------------------------- spatial conditional after the smoother (row f3)
// Posterior at new spatial points from the smoothed posterior (m_t, P_t) at the M inducing points, every time step
// independently (computation/spatial_conditionals.py:30-207 -> marginals.py:82-113, the `f_only` branch):
//   mu_t  = W m_t,                     W = Ksz Kzz^-1  [N x M]   (time-invariant, prepared by the caller)
//   var_t = ktt_t C0 + W (P_t + jitter I) W^T,   C0 = Kss - Ksz Kzz^-1 Kzs  [N x N]
// (the reference factors P_t + jitter I and multiplies the factor back: the sum above is what that evaluates to).
// One CTA per time step, persistent over t: U = jitter W^T + P_t W^T (M x N, per-CTA scratch that stays in L2), then
// either the full N x N block by a second tile GEMM on the lower triangle (mirrored) or only its diagonal.
struct SpatialCondArgs {
  int64_t T; int M, N;
  const double* W;        // [N, M]
  const double* Wt;       // [M, N]  W^T            (workspace, written by spatial_cond_prep_kernel)
  const double* jWt;      // [M, N]  jitter W^T     (workspace)
  const double* C0;       // [N, N]
  const double* ktt;      // [T] or null
  const double* m;        // [T, M]
  const double* P;        // [T, M, M]
  int diagonal;
  double* mu;             // [T, N]
  double* var;            // [T, N, N] or [T, N]
  double* scratch;        // per CTA: M N doubles
};

__global__ void spatial_cond_prep_kernel(const double* W, double* Wt, double* jWt, int M, int N, double jitter) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (int64_t)M * N;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx / N), i = (int)(idx - (int64_t)j * N);
    const double w = W[(int64_t)i * M + j];
    Wt[idx] = w;
    jWt[idx] = jitter * w;
  }
}

__global__ void __launch_bounds__(NTH, 1) kron_spatial_cond_kernel(const SpatialCondArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, M = p.M, N = p.N;
  double* U = p.scratch + (int64_t)blockIdx.x * M * N;
  for (int64_t t = blockIdx.x; t < p.T; t += gridDim.x) {
    const double* Pt = p.P + t * (int64_t)M * M;
    const double* mt = p.m + t * (int64_t)M;
    {
      Gemm g = {};
      g.npair = 1; g.nt = 1;
      g.A[0] = Pt; g.lda[0] = M; g.B[0] = p.W; g.ldb[0] = M; g.K[0] = M; g.sg[0] = 1.0;
      g.M = M; g.N = N;
      g.Add = p.jWt; g.ldadd = N;
      g.C = U; g.ldc = N;
      gemm_tiles<5>(g, 0, 1, sm);
    }
    // mean: one warp per new point
    for (int i = warp; i < N; i += NTH / 32) {
      double a = 0.0;
      for (int k = lane; k < M; k += 32) a = fma(__ldg(p.W + (int64_t)i * M + k), mt[k], a);
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) p.mu[t * N + i] = a;
    }
    __syncthreads();            // U complete (written by this CTA's own threads)
    if (p.diagonal) {
      const double kt = p.ktt ? p.ktt[t] : 1.0;
      for (int i = tid; i < N; i += NTH) {
        double a0 = 0.0, a1 = 0.0;
        int j = 0;
        for (; j + 2 <= M; j += 2) {
          a0 = fma(__ldg(p.Wt + (int64_t)j * N + i), __ldcg(U + (int64_t)j * N + i), a0);
          a1 = fma(__ldg(p.Wt + (int64_t)(j + 1) * N + i), __ldcg(U + (int64_t)(j + 1) * N + i), a1);
        }
        if (j < M) a0 = fma(__ldg(p.Wt + (int64_t)j * N + i), __ldcg(U + (int64_t)j * N + i), a0);
        p.var[t * N + i] = fma(kt, __ldg(p.C0 + (int64_t)i * N + i), a0 + a1);
      }
    } else {
      Gemm g = {};
      g.npair = 1; g.nt = 0;
      g.A[0] = p.W; g.lda[0] = M; g.B[0] = U; g.ldb[0] = N; g.K[0] = M; g.sg[0] = 1.0;
      g.M = N; g.N = N;
      g.Add = p.C0; g.ldadd = N; g.add_scale = p.ktt ? p.ktt + t : nullptr;
      g.C = p.var + t * (int64_t)N * N; g.ldc = N;
      g.lower = 1;
      gemm_tiles<4>(g, 0, 1, sm);
    }
    __syncthreads();            // U is overwritten by the next step
  }
}

// ------------------------------------------------------------------------------------------- filter
struct FilterArgs {
  int64_t T;
  int Ns, ds, d;
  const double* At; const double* Qt; const int32_t* idx;
  const double* Ks; const double* m0; const double* P0;
  const double* Y; const double* R; int64_t R_ts;
  double jitter;
  double* mf; double* Pf; double* lml;
  double* Pp; double* W; double* Kb; double* SjA; double* SmA; double* mp; double* acc;
  double* prof;
  unsigned long long* bar;
  int fuse_predict;
};

__global__ void __launch_bounds__(NTH, 1) kron_filter_kernel(const FilterArgs p) {
  extern __shared__ __align__(16) double sm[];
  GridBarrier grid;
  grid.init(p.bar);
  const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x;
  const int Ns = p.Ns, ds = p.ds, d = p.d, m = p.Ns;
  const int64_t gt = (int64_t)cta * NTH + tid, gstride = (int64_t)ncta * NTH;
  const int cta_lml = ncta > 1 ? 1 : 0;
  if (cta == cta_lml && tid == 0) p.acc[0] = 0.0;
  const int64_t dd = (int64_t)d * d, dm_ = (int64_t)d * m;
  // predict of step k: P_ = A P A^T + Q into Pp_o, innovation covariances, masked P_ H^T into W_o (all CTAs)
  auto predict_cov = [&](int64_t k, double* Pp_o, double* W_o) {
    const double* At = p.At + (int64_t)p.idx[k] * ds * ds;
    const double* Qt = p.Qt + (int64_t)p.idx[k] * ds * ds;
    const double* Pprev = k ? p.Pf + (k - 1) * dd : p.P0;
    const double* y = p.Y + k * m;
    const double* Rk = p.R + k * p.R_ts;
    double at[DSMAX][DSMAX], qt[DSMAX][DSMAX];
#pragma unroll
    for (int a = 0; a < DSMAX; ++a)
#pragma unroll
      for (int b = 0; b < DSMAX; ++b) {
        at[a][b] = (a < ds && b < ds) ? At[a * ds + b] : 0.0;
        qt[a][b] = (a < ds && b < ds) ? Qt[a * ds + b] : 0.0;
      }
    for (int64_t idx = gt; idx < (int64_t)Ns * Ns; idx += gstride) {
      const int I = (int)(idx / Ns), J = (int)(idx - (int64_t)I * Ns);
      double blk[DSMAX][DSMAX], tmp[DSMAX][DSMAX];
#pragma unroll
      for (int a = 0; a < DSMAX; ++a)
#pragma unroll
        for (int b = 0; b < DSMAX; ++b)
          blk[a][b] = (a < ds && b < ds) ? __ldcg(Pprev + (int64_t)(I * ds + a) * d + J * ds + b) : 0.0;
#pragma unroll
      for (int a = 0; a < DSMAX; ++a)
#pragma unroll
        for (int b = 0; b < DSMAX; ++b) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < DSMAX; ++c) s = fma(at[a][c], blk[c][b], s);
          tmp[a][b] = s;
        }
      const double ks = p.Ks[(int64_t)I * Ns + J];
      const double yi = y[I], yj = y[J];
      const bool oi = !(yi != yi), oj = !(yj != yj);
#pragma unroll
      for (int a = 0; a < DSMAX; ++a)
#pragma unroll
        for (int b = 0; b < DSMAX; ++b) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < DSMAX; ++c) s = fma(tmp[a][c], at[b][c], s);
          blk[a][b] = s + ks * qt[a][b];
          if (a < ds && b < ds) Pp_o[(int64_t)(I * ds + a) * d + J * ds + b] = blk[a][b];
        }
      const double s = ((oi && oj) ? blk[0][0] : 0.0) + Rk[(int64_t)I * Ns + J];
      p.SjA[(int64_t)I * m + J] = s + (I == J ? p.jitter : 0.0);
      p.SmA[(int64_t)I * m + J] = (oi && oj) ? s : (I == J ? 1.0 : 0.0);
#pragma unroll
      for (int a = 0; a < DSMAX; ++a)
        if (a < ds) W_o[(int64_t)(I * ds + a) * m + J] = oj ? blk[a][0] : 0.0;
    }
  };
  // predicted mean and innovation of step k (threads first, first + stride, ...)
  auto predict_mean = [&](int64_t k, int64_t first, int64_t stride) {
    const double* At = p.At + (int64_t)p.idx[k] * ds * ds;
    const double* mprev = k ? p.mf + (k - 1) * (int64_t)d : p.m0;
    const double* y = p.Y + k * m;
    for (int64_t I = first; I < Ns; I += stride) {
      double mu0 = 0.0;
      for (int a = 0; a < ds; ++a) {
        double s = 0.0;
        for (int b = 0; b < ds; ++b) s = fma(At[a * ds + b], __ldcg(mprev + I * ds + b), s);
        p.mp[I * ds + a] = s;
        if (a == 0) mu0 = s;
      }
      const double yi = y[I];
      const double v = (yi != yi) ? 0.0 : yi - mu0;
      p.SjA[(int64_t)m * m + I] = v;
      p.SmA[(int64_t)m * m + I] = v;
    }
  };
  // the predict of step k + 1 rides in the update epilogue of step k when tiles hold whole ds x ds blocks
  const bool fusable = (32 % ds) == 0 && p.fuse_predict;
  unsigned long long tick_ = gtimer();
  predict_cov(0, p.Pp, p.W);
  predict_mean(0, gt, gstride);
  KRON_TICK(0)
  grid.sync();
  KRON_TICK(1)
  for (int64_t k = 0; k < p.T; ++k) {
    double* Ppc = p.Pp + (k & 1) * dd;              // this step's predicted covariance and P_ H^T / X
    double* Wc = p.W + (k & 1) * dm_;
    double* Ppn = p.Pp + ((k + 1) & 1) * dd;        // the next step's
    double* Wn = p.W + ((k + 1) & 1) * dm_;
    const double* y = p.Y + k * m;
    // ---- phase 2: the two Cholesky factorisations, each with the innovation appended as an extra row
    const bool resident = m <= TRI_N;
    if (cta == 0) {
      if (resident) chol_smem(p.SjA, m, m, p.SjA, sm, p.prof);
      else chol_blocked(p.SjA, m, m, m + 1, sm, p.prof);
    }
    if (cta == cta_lml) {
      if (cta_lml == 0) __syncthreads();
      double ld = 0.0, mh = 0.0, no = 0.0;
      if (resident) {
        chol_smem(p.SmA, m, m, nullptr, sm, nullptr);
        for (int a = tid; a < m; a += NTH) {
          ld += log(sm[TS_L + tri_idx(a, a)]);
          const double u = sm[TS_V + a];
          mh = fma(u, u, mh);
        }
      } else {
        chol_blocked(p.SmA, m, m, m + 1, sm);
        for (int a = tid; a < m; a += NTH) {
          ld += log(__ldcg(p.SmA + (int64_t)a * m + a));
          const double u = __ldcg(p.SmA + (int64_t)m * m + a);
          mh = fma(u, u, mh);
        }
      }
      for (int a = tid; a < m; a += NTH) {
        const double ya = y[a];
        no += (ya != ya) ? 0.0 : 1.0;
      }
      __syncthreads();
      double* scr = sm + TS_T;
      ld = block_sum(ld, scr);
      mh = block_sum(mh, scr);
      no = block_sum(no, scr);
      if (tid == 0) p.acc[0] += -0.5 * (no * 1.8378770664093454835606594728112 + 2.0 * ld + mh);
    }
    KRON_TICK(2)
    grid.sync();
    KRON_TICK(3)
    // ---- phase 3: rows of X = W L^-T (in place in W), the mean update, rows of K = X L^-1
    for (int rb = cta; rb * 8 < d; rb += ncta) {
      const int i0 = rb * 8, R = min(8, d - i0);
      double* Xr = Wc + (int64_t)i0 * m;
      double* Kr = p.Kb + (int64_t)i0 * m;
      if (resident) {
        trsm_smem(p.SjA, m, m, Xr, R, p.jitter != 0.0 ? Kr : nullptr, p.SjA + (int64_t)m * m, p.mp + i0,
                  p.mf + k * (int64_t)d + i0, sm, p.prof);
        __syncthreads();
        continue;
      }
      trsm_fwd<1>(Xr, m, R, p.SjA, m, m, sm, p.prof);
      {
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < R) {
          double s = 0.0;
          for (int a = lane; a < m; a += 32) s = fma(__ldcg(Xr + (int64_t)warp * m + a), __ldcg(p.SjA + (int64_t)m * m + a), s);
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) p.mf[k * (int64_t)d + i0 + warp] = p.mp[i0 + warp] + s;
        }
      }
      if (p.jitter != 0.0) {
        for (int idx = tid; idx < R * m; idx += NTH) Kr[idx] = __ldcg(Xr + idx);
        __syncthreads();
        trsm_bwd<1>(Kr, m, R, p.SjA, m, m, nullptr, 0, sm);
      }
    }
    KRON_TICK(4)
    grid.sync();
    KRON_TICK(5)
    // ---- phase 4: P = P_ - X X^T + jitter K K^T  (== P_ - K S K^T with the un-jittered S), and -- fused into its
    // epilogue -- the predict of step k + 1
    const bool fuse = fusable && k + 1 < p.T;
    {
      PredictFuse fz;
      if (fuse) {
        fz.Ns = Ns; fz.ds = ds; fz.d = d;
        fz.At = p.At + (int64_t)p.idx[k + 1] * ds * ds; fz.Qt = p.Qt + (int64_t)p.idx[k + 1] * ds * ds;
        fz.Ks = p.Ks; fz.y = p.Y + (k + 1) * m; fz.R = p.R + (k + 1) * p.R_ts; fz.jitter = p.jitter;
        fz.Pp = Ppn; fz.W = Wn; fz.SjA = p.SjA; fz.SmA = p.SmA;
      }
      Gemm g = {};
      g.nt = 1;
      g.A[0] = Wc; g.lda[0] = m; g.B[0] = Wc; g.ldb[0] = m; g.K[0] = m; g.sg[0] = -1.0;
      g.npair = 1;
      if (p.jitter != 0.0) {
        g.A[1] = p.Kb; g.lda[1] = m; g.B[1] = p.Kb; g.ldb[1] = m; g.K[1] = m; g.sg[1] = p.jitter;
        g.npair = 2;
      }
      g.M = d; g.N = d;
      g.Add = Ppc; g.ldadd = d;
      g.C = p.Pf + k * dd; g.ldc = d;
      g.lower = 1;
      g.fuse = fuse ? &fz : nullptr;
      gemm_tiles<4>(g, cta, ncta, sm);
      if (fuse && cta == ncta - 1) predict_mean(k + 1, tid, NTH);      // mf[k] is complete since the last barrier
    }
    KRON_TICK(6)
    grid.sync();
    KRON_TICK(7)
    if (!fuse && k + 1 < p.T) {
      predict_cov(k + 1, Ppn, Wn);
      predict_mean(k + 1, gt, gstride);
      KRON_TICK(0)
      grid.sync();
      KRON_TICK(1)
    }
  }
  if (cta == cta_lml && tid == 0) p.lml[0] = p.acc[0];
}

// ------------------------------------------------------------------------------------------ smoother
struct GainArgs {
  int64_t k_lo, k_hi;       // steps of this time chunk (inclusive)
  int Ns, ds, d;
  const double* At; const double* Qt; const int32_t* idx;
  const double* Ks; const double* Pf;
  double jitter;
  double* Gc; double* Ppc;  // [k_hi - k_lo + 1, d, d] gains and (un-jittered) predicted covariances
  double* scratch;          // 2 d^2 doubles per CTA
};

// parallel over time: G_k = Pf_k A^T (A Pf_k A^T + Q + jitter I)^-1, one CTA per step
__global__ void __launch_bounds__(NTH, 1) kron_gain_kernel(const GainArgs p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, Ns = p.Ns, ds = p.ds, d = p.d;
  double* Wk = p.scratch + (int64_t)blockIdx.x * 2 * d * d;
  double* Lk = Wk + (int64_t)d * d;
  for (int64_t k = p.k_lo + blockIdx.x; k <= p.k_hi; k += gridDim.x) {
    const double* At = p.At + (int64_t)p.idx[k] * ds * ds;
    const double* Qt = p.Qt + (int64_t)p.idx[k] * ds * ds;
    const double* Pfk = p.Pf + k * (int64_t)d * d;
    double* Ppk = p.Ppc + (k - p.k_lo) * (int64_t)d * d;
    double at[DSMAX][DSMAX], qt[DSMAX][DSMAX];
#pragma unroll
    for (int a = 0; a < DSMAX; ++a)
#pragma unroll
      for (int b = 0; b < DSMAX; ++b) {
        at[a][b] = (a < ds && b < ds) ? At[a * ds + b] : 0.0;
        qt[a][b] = (a < ds && b < ds) ? Qt[a * ds + b] : 0.0;
      }
    for (int idx = tid; idx < Ns * Ns; idx += NTH) {
      const int I = idx / Ns, J = idx - I * Ns;
      double blk[DSMAX][DSMAX], wb[DSMAX][DSMAX];
#pragma unroll
      for (int a = 0; a < DSMAX; ++a)
#pragma unroll
        for (int b = 0; b < DSMAX; ++b)
          blk[a][b] = (a < ds && b < ds) ? __ldcg(Pfk + (int64_t)(I * ds + a) * d + J * ds + b) : 0.0;
#pragma unroll
      for (int a = 0; a < DSMAX; ++a)
#pragma unroll
        for (int b = 0; b < DSMAX; ++b) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < DSMAX; ++c) s = fma(blk[a][c], at[b][c], s);     // (Pf A^T) block
          wb[a][b] = s;
        }
      const double ks = p.Ks[(int64_t)I * Ns + J];
#pragma unroll
      for (int a = 0; a < DSMAX; ++a)
#pragma unroll
        for (int b = 0; b < DSMAX; ++b) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < DSMAX; ++c) s = fma(at[a][c], wb[c][b], s);
          s += ks * qt[a][b];
          if (a < ds && b < ds) {
            const int64_t o = (int64_t)(I * ds + a) * d + J * ds + b;
            Wk[o] = wb[a][b];
            Ppk[o] = s;
            Lk[o] = s + ((I == J && a == b) ? p.jitter : 0.0);
          }
        }
    }
    __syncthreads();
    chol_blocked(Lk, d, d, d, sm);
    trsm_fwd<5>(Wk, d, d, Lk, d, d, sm);
    trsm_bwd<5>(Wk, d, d, Lk, d, d, p.Gc + (k - p.k_lo) * (int64_t)d * d, d, sm);
    __syncthreads();
  }
}

struct RecArgs {
  int64_t k_lo, k_hi;
  int Ns, ds, d;
  const double* At; const int32_t* idx;
  const double* mf; const double* Pf;
  const double* Gc; const double* Ppc;
  int project;                 // 1: outputs are H ms [T, Ns], H Ps H^T [T, Ns, Ns]; 0: full state
  double* ms_out; double* Ps_out;
  double* ringP; double* ringm; double* dP; double* T1;
  double* prof;
  unsigned long long* bar;
};
__device__ __forceinline__ double* ps_state(const RecArgs& p, int64_t k) {
  return p.project ? p.ringP + (k & 1) * (int64_t)p.d * p.d : p.Ps_out + k * (int64_t)p.d * p.d;
}

// the recursion through time: P_s,k = Pf_k + G_k (P_s,k+1 - P_pred,k) G_k^T, m_s,k = mf_k + G_k (m_s,k+1 - A mf_k)
__global__ void __launch_bounds__(NTH, 1) kron_smooth_rec_kernel(const RecArgs p) {
  extern __shared__ __align__(16) double sm[];
  GridBarrier grid;
  grid.init(p.bar);
  const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x;
  const int Ns = p.Ns, ds = p.ds, d = p.d;
  const int64_t dd = (int64_t)d * d;
  const int64_t gt = (int64_t)cta * NTH + tid, gstride = (int64_t)ncta * NTH;
  {
    const double* Psn = ps_state(p, p.k_hi + 1);
    const double* Pp = p.Ppc + (p.k_hi - p.k_lo) * dd;
    for (int64_t i = gt; i < dd; i += gstride) p.dP[i] = __ldcg(Psn + i) - __ldcg(Pp + i);
  }
  grid.sync();
  unsigned long long tick_ = gtimer(), tcta_ = tick_;
  for (int64_t k = p.k_hi; k >= p.k_lo; --k) {
    const double* G = p.Gc + (k - p.k_lo) * dd;
    tcta_ = gtimer();
    {
      Gemm g = {};
      g.npair = 1; g.nt = 0;
      g.A[0] = G; g.lda[0] = d; g.B[0] = p.dP; g.ldb[0] = d; g.K[0] = d; g.sg[0] = 1.0;
      g.M = d; g.N = d;
      g.C = p.T1; g.ldc = d;
      g.prof = p.prof;
      gemm_tiles<5>(g, cta, ncta, sm);
    }
    KRON_TICK(8)
    if (p.prof && tid == 0) { const unsigned long long now_ = gtimer(); p.prof[32 + 2 * cta] += (double)(now_ - tcta_); tcta_ = now_; }
    {
      // mean: one warp per row, rows dealt round-robin over all warps of the grid (from the back, so the CTAs
      // that hold no second GEMM tile take them)
      const double* At = p.At + (int64_t)p.idx[k] * ds * ds;
      const double* mfk = p.mf + k * (int64_t)d;
      const double* msn = p.ringm + ((k + 1) & 1) * (int64_t)d;
      double* msk = p.ringm + (k & 1) * (int64_t)d;
      const int lane = tid & 31;
      // rows go to the CTAs that hold no tile of the product above when there are any (their work is then hidden
      // behind the GEMM), else to every CTA
      const int nt1 = ((d + 39) / 40) * ((d + 31) / 32);
      const int first = ncta > nt1 ? nt1 : 0;
      const int gw = cta >= first ? (cta - first) * (NTH / 32) + (tid >> 5) : d, nw = (ncta - first) * (NTH / 32);
      if (cta >= first) {
        double* dm = sm;                              // m_s,k+1 - A mf_k, once per CTA
        for (int c = tid; c < d; c += NTH) {
          const int J = c / ds, b = c - J * ds;
          double am = 0.0;
          for (int b2 = 0; b2 < ds; ++b2) am = fma(At[b * ds + b2], mfk[J * ds + b2], am);
          dm[c] = __ldcg(msn + c) - am;
        }
        __syncthreads();
        for (int row = gw; row < d; row += nw) {
          const double* gr = G + (int64_t)row * d;
          double s0 = 0.0, s1 = 0.0;
          int c = lane;
          for (; c + 32 < d; c += 64) {
            s0 = fma(__ldcg(gr + c), dm[c], s0);
            s1 = fma(__ldcg(gr + c + 32), dm[c + 32], s1);
          }
          if (c < d) s0 = fma(__ldcg(gr + c), dm[c], s0);
          double s = s0 + s1;
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) {
            const double v = mfk[row] + s;
            msk[row] = v;
            if (!p.project) p.ms_out[k * (int64_t)d + row] = v;
            else if (row % ds == 0) p.ms_out[k * (int64_t)Ns + row / ds] = v;
          }
        }
      }
    }
    KRON_TICK(9)
    grid.sync();
    KRON_TICK(10)
    tcta_ = gtimer();
    {
      Gemm g = {};
      g.npair = 1; g.nt = 1;
      g.A[0] = p.T1; g.lda[0] = d; g.B[0] = G; g.ldb[0] = d; g.K[0] = d; g.sg[0] = 1.0;
      g.M = d; g.N = d;
      g.Add = p.Pf + k * dd; g.ldadd = d;
      g.C = ps_state(p, k); g.ldc = d;
      g.lower = 1;
      if (k > p.k_lo) { g.C2 = p.dP; g.Sub2 = p.Ppc + (k - 1 - p.k_lo) * dd; g.ld2 = d; }
      if (p.project) { g.Cp = p.Ps_out + k * (int64_t)Ns * Ns; g.ldp = Ns; g.ds = ds; }
      gemm_tiles<4>(g, cta, ncta, sm);
    }
    KRON_TICK(11)
    if (p.prof && tid == 0) { const unsigned long long now_ = gtimer(); p.prof[32 + 2 * cta + 1] += (double)(now_ - tcta_); }
    grid.sync();
    KRON_TICK(12)
  }
}

// last step: smoothed == filtered
__global__ void kron_emit_last_kernel(const RecArgs p, int64_t T) {
  const int d = p.d, ds = p.ds, Ns = p.Ns;
  const int64_t dd = (int64_t)d * d, k = T - 1;
  const double* Pfk = p.Pf + k * dd;
  const double* mfk = p.mf + k * (int64_t)d;
  double* Pst = ps_state(p, k);
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = gt; i < dd; i += gs) {
    const double v = Pfk[i];
    Pst[i] = v;
    if (p.project) {
      const int r = (int)(i / d), c = (int)(i - (int64_t)r * d);
      if (r % ds == 0 && c % ds == 0) p.Ps_out[k * (int64_t)Ns * Ns + (int64_t)(r / ds) * Ns + c / ds] = v;
    }
  }
  for (int64_t i = gt; i < d; i += gs) {
    const double v = mfk[i];
    p.ringm[(k & 1) * (int64_t)d + i] = v;
    if (!p.project) p.ms_out[k * (int64_t)d + i] = v;
    else if (i % ds == 0) p.ms_out[k * (int64_t)Ns + i / ds] = v;
  }
}

// ----------------------------------------------------------------------------------------- host side
struct Dev {
  int sms = 0, coop = 0;
  int filter_blocks = 0, rec_blocks = 0, gain_blocks = 0;
};
int device_setup(Dev& dv) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_status(e, "kron: cudaGetDevice");
  cudaDeviceGetAttribute(&dv.sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&dv.coop, cudaDevAttrCooperativeLaunch, dev);
  if (!dv.coop) return set_error(PHYSS_ERR_UNSUPPORTED, "kron: device does not support cooperative launches");
  e = cudaFuncSetAttribute(kron_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES_FILTER);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kron_gain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kron_smooth_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES);
  if (e != cudaSuccess) return cuda_status(e, "kron: cudaFuncSetAttribute(shared memory)");
  int b = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kron_filter_kernel, NTH, SM_BYTES_FILTER);
  if (e != cudaSuccess || b < 1) return cuda_status(e == cudaSuccess ? cudaErrorLaunchOutOfResources : e, "kron: occupancy (filter)");
  dv.filter_blocks = dv.sms;            // one CTA per SM: the tile counts are sized for it
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kron_smooth_rec_kernel, NTH, SM_BYTES);
  if (e != cudaSuccess || b < 1) return cuda_status(e == cudaSuccess ? cudaErrorLaunchOutOfResources : e, "kron: occupancy (smoother)");
  dv.rec_blocks = dv.sms;
  dv.gain_blocks = dv.sms;
  return PHYSS_OK;
}

constexpr int64_t GAIN_CHUNK_WAVES = 4;   // time steps per smoother chunk = 4 x (CTAs of the gain kernel)

inline int64_t align2(int64_t n) { return (n + 1) & ~(int64_t)1; }

constexpr int64_t BAR_WORDS = 16 * 66;     // top counter + up to 65 groups of 16 CTAs, one 128-byte line each
struct FilterWs { int64_t Pp, W, Kb, SjA, SmA, mp, acc, prof, bar, total; };
FilterWs filter_ws(int Ns, int ds) {
  const int64_t d = (int64_t)Ns * ds, m = Ns;
  FilterWs w{};
  int64_t o = 0;
  w.Pp = o; o += 2 * align2(d * d);            // double-buffered: the predict of step k + 1 is written while the
  w.W = o; o += 2 * align2(d * m);             // tiles of step k are still being read
  w.Kb = o; o += align2(d * m);
  w.SjA = o; o += align2((m + 1) * m);
  w.SmA = o; o += align2((m + 1) * m);
  w.mp = o; o += align2(d);
  w.acc = o; o += 2;
  w.prof = o; o += 32;
  w.bar = o; o += BAR_WORDS;
  w.total = o;
  return w;
}
struct SmoothWs { int64_t Gc, Ppc, scratch, ringP, ringm, dP, T1, prof, bar, total, chunk; };
SmoothWs smooth_ws(int Ns, int ds, int64_t T, int gain_blocks) {
  const int64_t d = (int64_t)Ns * ds, dd = d * d;
  SmoothWs w{};
  w.chunk = GAIN_CHUNK_WAVES * gain_blocks;
  if (w.chunk > T - 1) w.chunk = T - 1 > 0 ? T - 1 : 1;
  int64_t o = 0;
  w.Gc = o; o += w.chunk * dd;
  w.Ppc = o; o += w.chunk * dd;
  w.scratch = o; o += (int64_t)gain_blocks * 2 * dd;
  w.ringP = o; o += 2 * dd;
  w.ringm = o; o += align2(2 * d);
  w.dP = o; o += dd;
  w.T1 = o; o += dd;
  w.prof = o; o += 32 + 2 * 1024;
  w.bar = o; o += BAR_WORDS;
  w.total = o;
  return w;
}

}  // namespace kron
}  // namespace physs

using namespace physs;
using namespace physs::kron;

extern "C" {

int64_t physs_kron_workspace_bytes(int64_t T, int32_t Ns, int32_t ds, int32_t smoother) {
  if (T < 1 || Ns < 1 || ds < 1 || ds > DSMAX) return 0;
  if (!smoother) return 8 * filter_ws(Ns, ds).total + 16;
  Dev dv;
  if (device_setup(dv) != PHYSS_OK) return 0;
  return 8 * smooth_ws(Ns, ds, T, dv.gain_blocks).total + 16;
}

/* measurement aid: offset (in doubles) of the 32 phase timers inside the workspace (PHYSS_KRON_PROF=1) */
int64_t physs_kron_prof_offset(int64_t T, int32_t Ns, int32_t ds, int32_t smoother) {
  if (!smoother) return filter_ws(Ns, ds).prof;
  Dev dv;
  if (device_setup(dv) != PHYSS_OK) return -1;
  return smooth_ws(Ns, ds, T, dv.gain_blocks).prof;
}

int physs_kf_filter_kron_f64(void* stream, int64_t T, int32_t Ns, int32_t ds, const double* At, const double* Qt,
                             const int32_t* idx, const double* Ks, const double* m0, const double* P0,
                             const double* Y, const double* R, int64_t R_tstride, double jitter, void* ws,
                             int64_t ws_bytes, double* mf, double* Pf, double* lml) {
  if (T < 1 || Ns < 1 || ds < 1 || ds > DSMAX) return set_error(PHYSS_ERR_BAD_ARG, "kron filter: bad sizes (1 <= ds <= 4)");
  if (!At || !Qt || !idx || !Ks || !m0 || !P0 || !Y || !R || !ws || !mf || !Pf || !lml)
    return set_error(PHYSS_ERR_BAD_ARG, "kron filter: null required pointer");
  if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return set_error(PHYSS_ERR_BAD_ARG, "kron filter: workspace must be 16-byte aligned");
  const FilterWs L = filter_ws(Ns, ds);
  if (ws_bytes < 8 * L.total) return set_error(PHYSS_ERR_BAD_ARG, "kron filter: workspace too small");
  Dev dv;
  if (int rc = device_setup(dv)) return rc;
  double* w = static_cast<double*>(ws);
  FilterArgs a{};
  a.T = T; a.Ns = Ns; a.ds = ds; a.d = Ns * ds;
  a.At = At; a.Qt = Qt; a.idx = idx; a.Ks = Ks; a.m0 = m0; a.P0 = P0; a.Y = Y; a.R = R; a.R_ts = R_tstride;
  a.jitter = jitter; a.mf = mf; a.Pf = Pf; a.lml = lml;
  a.fuse_predict = getenv("PHYSS_KRON_NOFUSE") ? 0 : 1;
  a.prof = getenv("PHYSS_KRON_PROF") ? w + L.prof : nullptr;
  if (a.prof) cudaMemsetAsync(a.prof, 0, 32 * sizeof(double), (cudaStream_t)stream);
  a.Pp = w + L.Pp; a.W = w + L.W; a.Kb = w + L.Kb; a.SjA = w + L.SjA; a.SmA = w + L.SmA; a.mp = w + L.mp; a.acc = w + L.acc;
  a.bar = reinterpret_cast<unsigned long long*>(w + L.bar);
  cudaError_t e = cudaMemsetAsync(a.bar, 0, BAR_WORDS * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_status(e, "kron filter: barrier reset");
  void* args[] = {&a};
  e = cudaLaunchCooperativeKernel((void*)kron_filter_kernel, dim3(dv.filter_blocks), dim3(NTH), args, SM_BYTES_FILTER,
                                  (cudaStream_t)stream);
  return cuda_status(e, "kron_filter_kernel launch");
}

int64_t physs_cvi_big_workspace_bytes(int32_t D) {
  if (D < 1 || D > TRI_N) return 0;
  Dev dv;
  if (device_setup(dv) != PHYSS_OK) return 0;
  return 8 * cvi_big_scratch(D) * dv.sms + 16;
}

static int cvi_big_launch(void* stream, bool update, CviBigArgs& a, void* ws, int64_t ws_bytes) {
  if (a.T < 0 || a.D < 1 || a.D > TRI_N) return set_error(PHYSS_ERR_UNSUPPORTED, "cvi (large blocks): 1 <= D <= 208");
  if (a.T == 0) return PHYSS_OK;
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 15) != 0) return set_error(PHYSS_ERR_BAD_ARG, "cvi (large blocks): workspace missing or not 16-byte aligned");
  Dev dv;
  if (int rc = device_setup(dv)) return rc;
  if (ws_bytes < 8 * cvi_big_scratch(a.D) * dv.sms) return set_error(PHYSS_ERR_BAD_ARG, "cvi (large blocks): workspace too small");
  a.scratch = static_cast<double*>(ws);
  const int grid = (int)(a.T < dv.sms ? a.T : dv.sms);
  cudaError_t e = cudaFuncSetAttribute(update ? (const void*)kron_cvi_site_kernel : (const void*)kron_cvi_ell_sur_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES_FILTER);
  if (e != cudaSuccess) return cuda_status(e, "cvi (large blocks): cudaFuncSetAttribute");
  if (update) kron_cvi_site_kernel<<<grid, NTH, SM_BYTES_FILTER, (cudaStream_t)stream>>>(a);
  else kron_cvi_ell_sur_kernel<<<grid, NTH, SM_BYTES_FILTER, (cudaStream_t)stream>>>(a);
  return cuda_status(cudaGetLastError(), "cvi (large blocks) launch");
}

int physs_cvi_natgrad_big_f64(void* stream, int64_t T, int32_t D, const double* Yt, const double* Vt, const double* qm,
                              const double* dm, const double* dS, int32_t dS_diag, double beta, double ngj, void* ws,
                              int64_t ws_bytes, double* Yn, double* Vn) {
  if (T > 0 && (!Yt || !Vt || !qm || !dm || !dS || !Yn || !Vn))
    return set_error(PHYSS_ERR_BAD_ARG, "cvi natgrad (large blocks): null required pointer");
  CviBigArgs a{};
  a.T = T; a.D = D; a.Yt = Yt; a.Vt = Vt; a.qm = qm; a.dm = dm; a.dS = dS; a.dS_diag = dS_diag ? 1 : 0;
  a.beta = beta; a.ngj = ngj; a.Yn = Yn; a.Vn = Vn;
  return cvi_big_launch(stream, true, a, ws, ws_bytes);
}

int physs_cvi_ell_sur_big_f64(void* stream, int64_t T, int32_t D, const double* Yt, const double* Vt, const double* qm,
                              const double* qS, void* ws, int64_t ws_bytes, double* ell) {
  if (T > 0 && (!Yt || !Vt || !qm || !qS || !ell))
    return set_error(PHYSS_ERR_BAD_ARG, "cvi surrogate ELL (large blocks): null required pointer");
  CviBigArgs a{};
  a.T = T; a.D = D; a.Yt = Yt; a.Vt = Vt; a.qm = qm; a.qS = qS; a.ell = ell;
  return cvi_big_launch(stream, false, a, ws, ws_bytes);
}

int64_t physs_spatial_conditional_ws_bytes(int32_t M, int32_t N) {
  if (M < 1 || N < 1) return 0;
  Dev dv;
  if (device_setup(dv) != PHYSS_OK) return 0;
  return 8 * ((int64_t)M * N * (dv.sms + 2)) + 16;
}

int physs_spatial_conditional_f64(void* stream, int64_t T, int32_t M, int32_t N, const double* W, const double* C0,
                                  const double* ktt, const double* m, const double* P, double jitter,
                                  int32_t diagonal, void* ws, int64_t ws_bytes, double* mu, double* var) {
  if (T < 0 || M < 1 || N < 1) return set_error(PHYSS_ERR_BAD_ARG, "spatial conditional: bad sizes");
  if (T == 0) return PHYSS_OK;
  if (!W || !C0 || !m || !P || !mu || !var) return set_error(PHYSS_ERR_BAD_ARG, "spatial conditional: null required pointer");
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 15) != 0)
    return set_error(PHYSS_ERR_BAD_ARG, "spatial conditional: workspace missing or not 16-byte aligned");
  Dev dv;
  if (int rc = device_setup(dv)) return rc;
  const int64_t MN = (int64_t)M * N;
  if (ws_bytes < 8 * MN * (dv.sms + 2)) return set_error(PHYSS_ERR_BAD_ARG, "spatial conditional: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* w = static_cast<double*>(ws);
  SpatialCondArgs a{};
  a.T = T; a.M = M; a.N = N; a.W = W; a.Wt = w; a.jWt = w + MN; a.scratch = w + 2 * MN;
  a.C0 = C0; a.ktt = ktt; a.m = m; a.P = P; a.diagonal = diagonal ? 1 : 0; a.mu = mu; a.var = var;
  spatial_cond_prep_kernel<<<(int)((MN + 255) / 256 < 1024 ? (MN + 255) / 256 : 1024), 256, 0, st>>>(W, w, w + MN, M, N, jitter);
  cudaError_t e = cudaFuncSetAttribute(kron_spatial_cond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES);
  if (e != cudaSuccess) return cuda_status(e, "spatial conditional: cudaFuncSetAttribute");
  const int grid = (int)(T < dv.sms ? T : dv.sms);
  kron_spatial_cond_kernel<<<grid, NTH, SM_BYTES, st>>>(a);
  return cuda_status(cudaGetLastError(), "spatial conditional launch");
}

int physs_rts_smooth_kron_f64(void* stream, int64_t T, int32_t Ns, int32_t ds, const double* At, const double* Qt,
                              const int32_t* idx, const double* Ks, const double* mf, const double* Pf,
                              int32_t project, double jitter, void* ws, int64_t ws_bytes, double* ms, double* Ps) {
  if (T < 1 || Ns < 1 || ds < 1 || ds > DSMAX) return set_error(PHYSS_ERR_BAD_ARG, "kron smoother: bad sizes (1 <= ds <= 4)");
  if (!At || !Qt || !idx || !Ks || !mf || !Pf || !ws || !ms || !Ps)
    return set_error(PHYSS_ERR_BAD_ARG, "kron smoother: null required pointer");
  if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return set_error(PHYSS_ERR_BAD_ARG, "kron smoother: workspace must be 16-byte aligned");
  Dev dv;
  if (int rc = device_setup(dv)) return rc;
  const SmoothWs L = smooth_ws(Ns, ds, T, dv.gain_blocks);
  if (ws_bytes < 8 * L.total) return set_error(PHYSS_ERR_BAD_ARG, "kron smoother: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* w = static_cast<double*>(ws);
  const int d = Ns * ds;
  RecArgs r{};
  r.Ns = Ns; r.ds = ds; r.d = d; r.At = At; r.idx = idx; r.mf = mf; r.Pf = Pf;
  r.Gc = w + L.Gc; r.Ppc = w + L.Ppc; r.project = project ? 1 : 0; r.ms_out = ms; r.Ps_out = Ps;
  r.ringP = w + L.ringP; r.ringm = w + L.ringm; r.dP = w + L.dP; r.T1 = w + L.T1;
  r.prof = getenv("PHYSS_KRON_PROF") ? w + L.prof : nullptr;
  if (r.prof) cudaMemsetAsync(r.prof, 0, (32 + 2 * 1024) * sizeof(double), st);
  kron_emit_last_kernel<<<dv.sms, 256, 0, st>>>(r, T);
  GainArgs g{};
  g.Ns = Ns; g.ds = ds; g.d = d; g.At = At; g.Qt = Qt; g.idx = idx; g.Ks = Ks; g.Pf = Pf; g.jitter = jitter;
  g.Gc = w + L.Gc; g.Ppc = w + L.Ppc; g.scratch = w + L.scratch;
  for (int64_t k_hi = T - 2; k_hi >= 0;) {
    const int64_t k_lo = k_hi - L.chunk + 1 > 0 ? k_hi - L.chunk + 1 : 0;
    g.k_lo = k_lo; g.k_hi = k_hi;
    const int64_t n = k_hi - k_lo + 1;
    const int gb = (int)(n < dv.gain_blocks ? n : dv.gain_blocks);
    kron_gain_kernel<<<gb, NTH, SM_BYTES, st>>>(g);
    r.k_lo = k_lo; r.k_hi = k_hi;
    r.bar = reinterpret_cast<unsigned long long*>(w + L.bar);
    cudaError_t e = cudaMemsetAsync(r.bar, 0, BAR_WORDS * sizeof(double), st);
    if (e != cudaSuccess) return cuda_status(e, "kron smoother: barrier reset");
    void* args[] = {&r};
    e = cudaLaunchCooperativeKernel((void*)kron_smooth_rec_kernel, dim3(dv.rec_blocks), dim3(NTH), args,
                                                SM_BYTES, st);
    if (e != cudaSuccess) return cuda_status(e, "kron_smooth_rec_kernel launch");
    k_hi = k_lo - 1;
  }
  return cuda_status(cudaGetLastError(), "kron smoother launches");
}

}  // extern "C"
