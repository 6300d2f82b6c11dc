// physs_seq_impl.cuh -- sequential Kalman filter / RTS smoother, one thread per series (d <= 4 fast path,
// d = 6, 8 interim), every d x d block in registers, compile-time unrolled.
//
// Reference call sites replaced: kalman_filter.py:439-485 (filter('sequential')) and
// rts_smoother.py:162-192 (smoother('sequential')); see include/physs_b200.h.
//
// Layout: every per-step array (Y, mf, Pf, ms, Ps, lml_k) is addressed through two "step strides":
// row (b, k) of an array with n doubles per step starts at base + (b * sbs + k * sts) * n.
//   (sbs, sts) = (T, 1): batch-major [B][T][n], the order jax.vmap(axis 0) of the reference gives;
//   (sbs, sts) = (1, B): time-major  [T][B][n], the layout the B200 path prefers -- the 32 series of a
//   warp then touch one contiguous 32 * n * 8-byte span per step, which the kernels write / read with
//   fully coalesced 16-byte accesses after a warp-local transpose through shared memory.
// One thread walks one series in time.  HBM traffic per state-step is the algorithmic
// 8*(d*d + d + m*m + m + 1) B for the filter and 8*2*(d*d + d) B for the smoother.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "physs_core.cuh"
#include "physs_internal.h"

namespace physs {

template <int N>
__device__ __forceinline__ void load_vec(const double* __restrict__ src, double (&dst)[N]) {
  // every base pointer of the ABI is 16-byte aligned (checked in physs_api.cu), so rows with an even
  // number of doubles are always 16-byte aligned
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const double2 v = reinterpret_cast<const double2*>(src)[i];
      dst[2 * i] = v.x;
      dst[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = src[i];
  }
}

template <int N>
__device__ __forceinline__ void store_vec(double* __restrict__ dst, const double (&src)[N]) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i)
      reinterpret_cast<double2*>(dst)[i] = make_double2(src[2 * i], src[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = src[i];
  }
}

template <int D>
__device__ __forceinline__ void load_mat(const double* __restrict__ src, double (&dst)[D][D]) {
  load_vec<D * D>(src, *reinterpret_cast<double (*)[D * D]>(&dst[0][0]));
}
template <int D>
__device__ __forceinline__ void store_mat(double* __restrict__ dst, const double (&src)[D][D]) {
  store_vec<D * D>(dst, *reinterpret_cast<const double (*)[D * D]>(&src[0][0]));
}

template <int D, int S>
__device__ __forceinline__ void load_trans_dense(const double* __restrict__ src, Trans<D, S>& A) {
  static_assert(D == S, "dense transition is a single block");
  load_vec<D * D>(src, *reinterpret_cast<double (*)[D * D]>(&A.a[0][0][0]));
}

template <int D, int S>
__device__ __forceinline__ void matern_trans(const double (&lam)[D / S], double dt, Trans<D, S>& A) {
#pragma unroll
  for (int b = 0; b < D / S; ++b) MaternExpm<S>::eval(lam[b], dt, A.a[b]);
}

// integrated Wiener blocks: A_k and Q_k of every block in closed form (kernels/wiener.py:105-149), Q block-diagonal
template <int D, int S>
__device__ __forceinline__ void iwp_trans(const double (&var)[D / S], double dt, Trans<D, S>& A, double (&Q)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) Q[i][j] = 0.0;
  }
#pragma unroll
  for (int b = 0; b < D / S; ++b) {
    double q[S][S];
    IwpDisc<S>::eval(var[b], dt, A.a[b], q);
#pragma unroll
    for (int i = 0; i < S; ++i) {
#pragma unroll
      for (int j = 0; j < S; ++j) Q[b * S + i][b * S + j] = q[i][j];
    }
  }
}

// ------------------------------------------------------------------- warp-cooperative row transfer
// Rows of N doubles owned one per lane, 32 consecutive series contiguous in global memory (sbs == 1).
// LD = padded row length of the shared-memory tile.
// Even N moves 16-byte pieces: a quarter warp (8 lanes) is conflict-free when the row stride is an ODD number of
// 16-byte units, i.e. LD / 2 odd -- N itself when N / 2 is odd (N = 2, 6, 10, 14: the packed rows), else N + 2.
template <int N>
struct RowTile {
  static constexpr int LD = (N % 2 == 0) ? (((N / 2) % 2 == 1) ? N : N + 2) : N;
  static constexpr int SIZE = 32 * LD;
};

// lane `lane` contributes v[N]; the warp writes rows [0, nvalid) to g (row r at g + r * N)
template <int N>
__device__ __forceinline__ void warp_store_rows(double* __restrict__ g, const double (&v)[N],
                                                double* tile, int lane, int nvalid) {
  constexpr int LD = RowTile<N>::LD;
  if constexpr (N % 2 == 0) {
#pragma unroll
    for (int j = 0; j < N / 2; ++j)
      *reinterpret_cast<double2*>(tile + lane * LD + 2 * j) = make_double2(v[2 * j], v[2 * j + 1]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const int e = 2 * (i * 32 + lane);
      const int r = e / N, c = e % N;
      const double2 t = *reinterpret_cast<const double2*>(tile + r * LD + c);
      if (r < nvalid) *reinterpret_cast<double2*>(g + e) = t;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) tile[lane * LD + j] = v[j];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int e = i * 32 + lane;
      const int r = e / N, c = e % N;
      if (r < nvalid) g[e] = tile[r * LD + c];
    }
  }
  __syncwarp();
}

// phase 1 of a coalesced load: issue the global loads (rows >= nvalid shadow row nvalid - 1)
template <int N>
struct RowRaw {
  double2 v2[(N % 2 == 0) ? N / 2 : 1];
  double v1[(N % 2 == 0) ? 1 : N];
};
template <int N>
__device__ __forceinline__ void warp_load_issue(const double* __restrict__ g, RowRaw<N>& raw, int lane,
                                                int nvalid) {
  if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const int e = 2 * (i * 32 + lane);
      const int r = e / N, c = e % N;
      const int rr = (r < nvalid) ? r : nvalid - 1;
      raw.v2[i] = *reinterpret_cast<const double2*>(g + rr * N + c);
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int e = i * 32 + lane;
      const int r = e / N, c = e % N;
      const int rr = (r < nvalid) ? r : nvalid - 1;
      raw.v1[i] = g[rr * N + c];
    }
  }
}
// phase 2: transpose through the tile into the lane's own row
template <int N>
__device__ __forceinline__ void warp_load_finish(const RowRaw<N>& raw, double (&v)[N], double* tile,
                                                 int lane) {
  constexpr int LD = RowTile<N>::LD;
  if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const int e = 2 * (i * 32 + lane);
      *reinterpret_cast<double2*>(tile + (e / N) * LD + (e % N)) = raw.v2[i];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      const double2 t = *reinterpret_cast<const double2*>(tile + lane * LD + 2 * j);
      v[2 * j] = t.x;
      v[2 * j + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int e = i * 32 + lane;
      tile[(e / N) * LD + (e % N)] = raw.v1[i];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = tile[lane * LD + j];
  }
  __syncwarp();
}

template <int D>
__device__ __forceinline__ double (&flat(double (&P)[D][D]))[D * D] {
  return *reinterpret_cast<double (*)[D * D]>(&P[0][0]);
}
template <int D>
__device__ __forceinline__ const double (&flat(const double (&P)[D][D]))[D * D] {
  return *reinterpret_cast<const double (*)[D * D]>(&P[0][0]);
}

// Packed hand-over row of the fused filter + smoother call: [m (D) | upper triangle of P, row-major (D (D + 1) / 2)]
// padded to an even number of doubles (16-byte pieces).  The filter's update mirrors P, so the row carries all of it.
template <int D>
struct PackedRow {
  static constexpr int N = (D + D * (D + 1) / 2 + 1) & ~1;
};
template <int D>
__device__ __forceinline__ void pack_row(const double (&m)[D], const double (&P)[D][D], double (&row)[PackedRow<D>::N]) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) row[n++] = m[i];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = i; j < D; ++j) row[n++] = P[i][j];
  }
#pragma unroll
  for (int i = D + D * (D + 1) / 2; i < PackedRow<D>::N; ++i) row[i] = 0.0;
}
template <int D>
__device__ __forceinline__ void unpack_row(const double (&row)[PackedRow<D>::N], double (&m)[D], double (&P)[D][D]) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) m[i] = row[n++];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = i; j < D; ++j) {
      P[i][j] = row[n];
      P[j][i] = row[n++];
    }
  }
}
// the filter's transpose tile holds the larger of a P row and a packed row
template <int D>
struct SeqTile {
  static constexpr int SIZE = (RowTile<D * D>::SIZE > RowTile<PackedRow<D>::N>::SIZE) ? RowTile<D * D>::SIZE
                                                                                       : RowTile<PackedRow<D>::N>::SIZE;
};

// warps per block: the row tile of the largest per-step block must fit the 48 KB static shared memory
template <int D>
struct SeqBlock {
  static constexpr int WARPS = (RowTile<D * D>::SIZE * 8 * 4 <= 48 * 1024) ? 4 : 2;
  static constexpr int THREADS = 32 * WARPS;
};

// (series, chunk) of a thread.  Series are padded to a multiple of 32 per chunk so that a warp never
// straddles two chunks; lanes beyond B shadow series B - 1 and never store.
struct SeqWork {
  int64_t b, b0, c, v, t0, T;   // b0 = first series of this warp
  bool active;
  int nvalid;   // valid lanes of this warp
};
template <bool CHUNK, typename Args>
__device__ __forceinline__ bool seq_work(const Args& p, SeqWork& w) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t Bp = (p.B + 31) & ~(int64_t)31;
  const int64_t ci = tid / Bp;
  if (ci >= (CHUNK ? p.chunk_count : 1)) return false;      // whole warps only
  const int64_t bq = tid % Bp;
  w.active = bq < p.B;
  w.b = w.active ? bq : p.B - 1;
  const int64_t warp0 = bq - (threadIdx.x & 31);
  w.b0 = warp0;
  const int64_t left = p.B - warp0;
  w.nvalid = left >= 32 ? 32 : (int)left;
  w.c = CHUNK ? p.chunk_first + ci : 0;
  w.v = CHUNK ? w.b * p.nchunk + w.c : w.b;
  w.t0 = CHUNK ? w.c * p.chunk_len : 0;
  w.T = CHUNK ? ((p.chunk_len < p.T - w.t0) ? p.chunk_len : (p.T - w.t0)) : p.T;
  return true;
}

// ------------------------------------------------------------------------------------------ filter
// relative agreement of a freshly computed (m, P) with the stored one (fix-up pass of chunk mode)
template <int D>
__device__ __forceinline__ bool agrees(const double (&m)[D], const double (&P)[D][D],
                                       const double* __restrict__ om, const double* __restrict__ oP,
                                       double delta) {
  double dP = 0.0, sP = 0.0, dm = 0.0, sm = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const double omi = om[i];
    dm = fmax(dm, fabs(m[i] - omi));
    sm = fmax(sm, fabs(omi));
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double o = oP[i * D + j];
      dP = fmax(dP, fabs(P[i][j] - o));
      sP = fmax(sP, fabs(o));
    }
  }
  // mean: relative to its own magnitude, or to the posterior standard deviation if that is larger
  return (dP <= delta * sP) && (dm <= delta * sm || dm * dm <= delta * delta * sP);
}

// PACK (plain mode, time-major steps, D <= 4): the filtered moments go out as packed rows [m | triu(P)] (SeqFilterArgs::pk)
// instead of (mf, Pf).  A template argument, not a run-time branch: the branch cost the unpacked kernel 14 % of its
// cycles (12.5 -> 13.5 ms per 32,768 x 10k; measured at the end of round 2).
template <int D, int S, int M, bool HID, int GIVEN, bool CHUNK, bool PACK = false>
__global__ void __launch_bounds__(SeqBlock<D>::THREADS) seq_filter_kernel(const SeqFilterArgs p) {
  __shared__ __align__(16) double tiles[SeqBlock<D>::WARPS][PACK ? SeqTile<D>::SIZE : RowTile<D * D>::SIZE];
  SeqWork wk;
  if (!seq_work<CHUNK>(p, wk)) return;
  if (CHUNK && p.fixup && p.prev_changed && *p.prev_changed == 0) return;   // the previous pass was a fixed point
  constexpr int NB = D / S;
  const int64_t b = wk.b, v = wk.v, t0 = wk.t0, T = wk.T;
  const bool active = wk.active;
  const int lane = threadIdx.x & 31;
  double* tile = tiles[threadIdx.x >> 5];
  const bool coal = (p.sbs == 1);
  const int64_t sts = p.sts;

  // GIVEN: 0 = closed-form Matern blocks (lam, Pinf), 1 = A_k, Q_k supplied, 2 = integrated Wiener blocks
  // (lam = spectral density, no stationary covariance)
  double m[D], P[D][D], Pinf[D][D], H[M][D], lam[NB];
  if (CHUNK && p.from_bnd) {
    load_vec<D>(p.bnd_m + v * D, m);
    load_mat<D>(p.bnd_P + v * D * D, P);
  } else {
    load_vec<D>(p.m0 + b * p.m0_bs, m);
    load_mat<D>(p.P0 + b * p.P0_bs, P);
  }
  if (GIVEN == 0) load_mat<D>(p.Pinf + b * p.Pinf_bs, Pinf);
  if (GIVEN != 1) {
#pragma unroll
    for (int i = 0; i < NB; ++i) lam[i] = p.lam[b * p.lam_bs + i];
  }
  if (!HID) {
#pragma unroll
    for (int a = 0; a < M; ++a) {
#pragma unroll
      for (int j = 0; j < D; ++j) H[a][j] = p.H[b * p.H_bs + a * D + j];
    }
  }
  const int64_t row0 = b * p.sbs + t0 * sts;                  // step-row index of (b, t0)
  const int64_t wrow0 = wk.b0 * p.sbs + t0 * sts;             // ... of the warp's first lane
  const double* __restrict__ dtp = p.dt + b * p.dt_bs + t0;
  const double* __restrict__ Yp = p.Y + row0 * M;
  const double* __restrict__ Rp = p.R + b * p.R_bs + t0 * p.R_ts;
  const double* __restrict__ Ap = (GIVEN == 1) ? p.A + b * p.A_bs + t0 * D * D : nullptr;
  const double* __restrict__ Qp = (GIVEN == 1) ? p.Q + b * p.Q_bs + t0 * D * D : nullptr;
  double* __restrict__ mfp = p.mf + row0 * D;
  double* __restrict__ Pfp = p.Pf + row0 * D * D;
  double* __restrict__ mfw = p.mf + wrow0 * D;
  double* __restrict__ Pfw = p.Pf + wrow0 * D * D;
  double* __restrict__ lkp = p.lml_k ? p.lml_k + row0 : nullptr;
  int streak = 0;
  bool done = false;

  LmlAcc acc;
  // speculative chunk mode: start `warm` steps before the chunk from (m0, P0) and discard those steps
  // (uniform over the warp: its lanes share the chunk)
  const int64_t w0 = (CHUNK && !p.from_bnd && p.warm > 0) ? ((p.warm < t0) ? p.warm : t0) : 0;
  // software prefetch of the next step's streamed inputs; the closed-form transition of step k + 1 is evaluated
  // during step k (it depends on dt only, so its exp / polynomial chain overlaps the predict / update chain of
  // the current step inside one basic block instead of heading the critical path of the next one)
  double y_n[M], R_n[M][M], dt_n;
  load_vec<M>(Yp - w0 * sts * M, y_n);
  load_mat<M>(Rp - w0 * p.R_ts, R_n);
  dt_n = (1 - w0 < T) ? dtp[1 - w0] : 0.0;                   // dt of the step AFTER the first one
  Trans<D, S> A_n;
  if constexpr (GIVEN == 0) matern_trans<D, S>(lam, dtp[-w0], A_n);
  double dt_cur = dtp[-w0];                                  // integrated Wiener: dt of the current step

  for (int64_t k = -w0; k < T; ++k) {
    double y[M], R[M][M];
    const double dt_next = dt_n;
#pragma unroll
    for (int a = 0; a < M; ++a) {
      y[a] = y_n[a];
#pragma unroll
      for (int c = 0; c < M; ++c) R[a][c] = R_n[a][c];
    }
    if (k + 1 < T) {
      load_vec<M>(Yp + (k + 1) * sts * M, y_n);
      if (p.R_ts != 0) load_mat<M>(Rp + (k + 1) * p.R_ts, R_n);
    }
    if (k + 2 < T) dt_n = dtp[k + 2];
    Trans<D, S> A;
    if constexpr (GIVEN == 1) {
      double Q[D][D];
      load_trans_dense<D, S>(Ap + k * D * D, A);
      load_mat<D>(Qp + k * D * D, Q);
      kf_predict_givenQ<D, S>(A, Q, m, P);
    } else if constexpr (GIVEN == 2) {
      double Q[D][D];
      iwp_trans<D, S>(lam, dt_cur, A, Q);
      dt_cur = dt_next;
      kf_predict_givenQ<D, S>(A, Q, m, P);
    } else {
      A = A_n;
      matern_trans<D, S>(lam, dt_next, A_n);                  // unconditional (dt_next is stale past the end: unused)
      kf_predict_stationary<D, S>(A, Pinf, m, P);
    }
    double det, mahal;
    int nobs;
    kf_update<D, M, HID>(m, P, H, R, y, p.jitter, det, mahal, nobs);
    acc.add(det, mahal, nobs);
    if (CHUNK && p.fixup && !done) {
      streak = agrees<D>(m, P, mfp + k * sts * D, Pfp + k * sts * D * D, p.delta) ? streak + 1 : 0;
      if (streak == 0 && active && p.pass_changed) atomicOr(p.pass_changed, 1);
    }
    if (k < 0) continue;                                   // warm-up step: nothing is stored
    if constexpr (PACK) {
      // packed hand-over to the smoother of the same call: one row [m | triu(P)] per series-step
      constexpr int NP = PackedRow<D>::N;
      double row[NP];
      pack_row<D>(m, P, row);
      warp_store_rows<NP>(p.pk + (wrow0 + k * sts) * NP, row, tile, lane, wk.nvalid);
    } else if (coal) {
      // converged lanes keep rewriting what is already stored to `delta`; the warp leaves together
      warp_store_rows<D>(mfw + k * sts * D, m, tile, lane, wk.nvalid);
      warp_store_rows<D * D>(Pfw + k * sts * D * D, flat<D>(P), tile, lane, wk.nvalid);
    } else if (active && !done) {
      store_vec<D>(mfp + k * sts * D, m);
      store_mat<D>(Pfp + k * sts * D * D, P);
    }
    if (lkp && active && !done) lkp[k * sts] = lml_term(det, mahal, nobs);
    if (CHUNK && p.fixup) {
      if (streak >= p.patience) done = true;
      if (__all_sync(0xffffffffu, done || !active)) break;
    }
  }
  if (CHUNK) {
    if (p.fixup && active && !done) atomicOr(p.unconverged, 1);
  } else if (active) {
    p.lml[b] = acc.value();
  }
}

// ---------------------------------------------------------------------------------------- smoother
// MO == 0: full_state (H = I).  MO > 0: project with Hout [MO, D].
template <int D, int S, int MO, int GIVEN, bool CHUNK, bool COAL>
__global__ void __launch_bounds__(SeqBlock<D>::THREADS) seq_smooth_kernel(const SeqSmoothArgs p) {
  __shared__ __align__(16) double tiles[SeqBlock<D>::WARPS][RowTile<D * D>::SIZE];
  SeqWork wk;
  if (!seq_work<CHUNK>(p, wk)) return;
  constexpr int NB = D / S;
  constexpr int MP = (MO == 0) ? D : MO;
  const int64_t b = wk.b, v = wk.v, t0 = wk.t0, T = wk.T;
  const bool active = wk.active;
  const int lane = threadIdx.x & 31;
  double* tile = tiles[threadIdx.x >> 5];
  constexpr bool coal = COAL;
  const int64_t sts = p.sts;

  double Pinf[D][D], lam[NB], Ho[MP][D];
  if (GIVEN == 0) load_mat<D>(p.Pinf + b * p.Pinf_bs, Pinf);
  if (GIVEN != 1) {
#pragma unroll
    for (int i = 0; i < NB; ++i) lam[i] = p.lam[b * p.lam_bs + i];
  }
  if (MO != 0) {
#pragma unroll
    for (int a = 0; a < MP; ++a) {
#pragma unroll
      for (int j = 0; j < D; ++j) Ho[a][j] = p.Hout[a * D + j];
    }
  }
  const int64_t row0 = b * p.sbs + t0 * sts;
  const int64_t wrow0 = wk.b0 * p.sbs + t0 * sts;
  const double* __restrict__ dtp = p.dt + b * p.dt_bs + t0;
  const double* __restrict__ Ap = (GIVEN == 1) ? p.A + b * p.A_bs + t0 * D * D : nullptr;
  const double* __restrict__ Qp = (GIVEN == 1) ? p.Q + b * p.Q_bs + t0 * D * D : nullptr;
  const double* __restrict__ mfp = p.mf + row0 * D;
  const double* __restrict__ Pfp = p.Pf + row0 * D * D;
  const double* __restrict__ mfw = p.mf + wrow0 * D;
  const double* __restrict__ Pfw = p.Pf + wrow0 * D * D;
  double* __restrict__ msp = p.ms + row0 * MP;
  double* __restrict__ Psp = p.Ps + row0 * MP * MP;
  double* __restrict__ msw = p.ms + wrow0 * MP;
  double* __restrict__ Psw = p.Ps + wrow0 * MP * MP;
  bool done = false;

  auto emit = [&](int64_t k, const double (&ms)[D], const double (&Ps)[D][D]) {
    if constexpr (MO == 0) {
      if constexpr (coal) {
        warp_store_rows<D>(msw + k * sts * D, ms, tile, lane, wk.nvalid);
        warp_store_rows<D * D>(Psw + k * sts * D * D, flat<D>(Ps), tile, lane, wk.nvalid);
      } else if (active && !done) {
        store_vec<D>(msp + k * sts * D, ms);
        store_mat<D>(Psp + k * sts * D * D, Ps);
      }
    } else {
      double om[MP], oP[MP][MP], HPs[MP][D];
#pragma unroll
      for (int a = 0; a < MP; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(Ho[a][j], ms[j], acc);
        om[a] = acc;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
#pragma unroll
          for (int c = 0; c < D; ++c) t = fma(Ho[a][c], Ps[c][j], t);
          HPs[a][j] = t;
        }
      }
#pragma unroll
      for (int a = 0; a < MP; ++a) {
#pragma unroll
        for (int c = 0; c < MP; ++c) {
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) t = fma(HPs[a][j], Ho[c][j], t);
          oP[a][c] = t;
        }
      }
      if (active) {
        store_vec<MP>(msp + k * sts * MP, om);
        store_mat<MP>(Psp + k * sts * MP * MP, oP);
      }
    }
  };

  // filtered moments of step k: direct per-thread loads, or (time-major) coalesced warp loads that are
  // transposed through shared memory when they are consumed one iteration later
  RowRaw<COAL ? D : 1> raw_m;
  RowRaw<COAL ? D * D : 1> raw_P;
  double mf_n[COAL ? 1 : D], Pf_n[COAL ? 1 : D][COAL ? 1 : D];
  auto fetch = [&](int64_t k) {
    if constexpr (coal) {
      warp_load_issue<D>(mfw + k * sts * D, raw_m, lane, wk.nvalid);
      warp_load_issue<D * D>(Pfw + k * sts * D * D, raw_P, lane, wk.nvalid);
    } else {
      load_vec<D>(mfp + k * sts * D, mf_n);
      load_mat<D>(Pfp + k * sts * D * D, Pf_n);
    }
  };
  auto land = [&](double (&mf)[D], double (&Pf)[D][D]) {
    if constexpr (coal) {
      warp_load_finish<D>(raw_m, mf, tile, lane);
      warp_load_finish<D * D>(raw_P, flat<D>(Pf), tile, lane);
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        mf[i] = mf_n[i];
#pragma unroll
        for (int j = 0; j < D; ++j) Pf[i][j] = Pf_n[i][j];
      }
    }
  };

  double ms[D], Ps[D][D];
  // plain mode: the last step is terminal (smoothed = filtered).  Chunk mode: every step of the chunk
  // is an RTS step from the carried state of the next chunk's first step; the very last chunk carries
  // its own last filtered state across dt = 0, which reproduces the terminal condition.
  // speculative chunk mode (warm > 0): no boundary is known; start `w0` steps past the chunk's end from the
  // FILTERED state there (a terminal-like condition) and discard the steps outside the chunk
  const bool spec = CHUNK && p.warm > 0;
  const int64_t after = p.T - (t0 + T);
  const int64_t w0 = spec ? ((p.warm < after) ? p.warm : after) : 0;
  const bool carried = CHUNK && !spec && (wk.c < p.nchunk - 1 || p.carry_last);
  int64_t kstart;
  if (carried) {
    load_vec<D>(p.bnd_m + v * D, ms);
    load_mat<D>(p.bnd_P + v * D * D, Ps);
  } else {
    load_vec<D>(mfp + (T - 1 + w0) * sts * D, ms);
    load_mat<D>(Pfp + (T - 1 + w0) * sts * D * D, Ps);
  }
  if (CHUNK) {
    kstart = (w0 > 0) ? T - 2 + w0 : T - 1;
  } else {
    emit(T - 1, ms, Ps);
    kstart = T - 2;
  }
  int streak = 0;

  double dt_n = 0.0;
  if (kstart >= 0) {
    fetch(kstart);
    dt_n = dtp[kstart];
  }
  for (int64_t k = kstart; k >= 0; --k) {
    double mf[D], Pf[D][D];
    const double dt = dt_n;
    land(mf, Pf);
    if (k >= 1) {
      fetch(k - 1);
      dt_n = dtp[k - 1];
    }
    Trans<D, S> A;
    if constexpr (GIVEN == 1) {
      double Q[D][D];
      load_trans_dense<D, S>(Ap + k * D * D, A);
      load_mat<D>(Qp + k * D * D, Q);
      rts_step<D, S>(A, Q, false, mf, Pf, p.jitter, ms, Ps);
    } else if constexpr (GIVEN == 2) {
      double Q[D][D];
      iwp_trans<D, S>(lam, dt, A, Q);
      rts_step<D, S>(A, Q, false, mf, Pf, p.jitter, ms, Ps);
    } else {
      matern_trans<D, S>(lam, dt, A);
      rts_step<D, S>(A, Pinf, true, mf, Pf, p.jitter, ms, Ps);
    }
    if constexpr (MO == 0) {      // the fix-up comparison needs the stored full state
      if (CHUNK && p.fixup && !done) {
        streak = agrees<D>(ms, Ps, msp + k * sts * D, Psp + k * sts * D * D, p.delta) ? streak + 1 : 0;
      }
    }
    if (k < T) emit(k, ms, Ps);                             // steps past the chunk's end are warm-up
    if (CHUNK && p.fixup) {
      if (streak >= p.patience) done = true;
      if (__all_sync(0xffffffffu, done || !active)) break;
    }
  }
  if (CHUNK && p.fixup && active && !done) atomicOr(p.unconverged, 1);
}

// ------------------------------------------------------------- software-pipelined smoother (time-major)
// The RTS step of series b at step k splits into a FRONT half that needs only the filtered moments of step k
// (A_k, m_pred, P_pred, Cholesky, gain G_k: ~3/4 of the arithmetic and all of its long dependency chains) and a
// BACK half (ms_k = mf_k + G_k (ms_{k+1} - m_pred), Ps_k = Pf_k + G_k (Ps_{k+1} - P_pred) G_k^T) that is the only
// part chained through time.  One thread still walks one series, but every loop iteration now holds the front of
// step k and the back of step k + 1 in ONE basic block with no dependence between them, so the instruction
// scheduler overlaps the Cholesky / substitution chains of one with the matrix products of the other (the batch
// leaves < 2 warps per scheduler: instruction-level parallelism is the only latency hiding there is).
// The filtered rows arrive through a 3-stage cp.async ring per warp (16-byte LDGSTS straight into the padded
// transpose tile, no staging registers), issued one iteration ahead; the back half re-reads (mf, Pf) of its step
// from the ring instead of carrying them in registers.  Same arithmetic, operation for operation, as rts_step.
__device__ __forceinline__ void seq_cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void seq_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void seq_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows [0, 32) of N doubles, contiguous in global memory at g, into the padded tile (row r at tile + r * LD)
template <int N>
__device__ __forceinline__ void warp_rows_async(const double* __restrict__ g, double* tile, int lane, int nvalid) {
  static_assert(N % 2 == 0, "16-byte pieces");
  constexpr int LD = RowTile<N>::LD;
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const int e = 2 * (i * 32 + lane);
    const int r = e / N, c = e % N;
    const int rr = (r < nvalid) ? r : nvalid - 1;
    seq_cp_async16(tile + r * LD + c, g + rr * N + c);
  }
}
template <int N>
__device__ __forceinline__ void tile_own_row(const double* tile, int lane, double (&v)[N]) {
  constexpr int LD = RowTile<N>::LD;
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    const double2 t = *reinterpret_cast<const double2*>(tile + lane * LD + 2 * j);
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}

template <int D>
struct SeqPipe {
  static constexpr int NST = 3;                                            // ring stages
  static constexpr int STAGE = RowTile<D>::SIZE + RowTile<D * D>::SIZE;    // doubles per stage (mf | Pf)
  static constexpr int PER_WARP = NST * STAGE + RowTile<D * D>::SIZE;      // + the output transpose tile
  static constexpr size_t smem_bytes(int warps) { return (size_t)warps * PER_WARP * sizeof(double); }
};

// PACK: the filtered moments come as packed rows [m | triu(P)] (SeqSmoothArgs::pk, written by the filter of the same
// physs_kf_filter_smooth_packed_f64 call) -- one ring tile of PackedRow<D>::N doubles per stage instead of two.
template <int D, int S, int MO, int GIVEN, bool PACK = false>
__global__ void __launch_bounds__(128) seq_smooth_pipe_kernel(const SeqSmoothArgs p) {
  extern __shared__ __align__(16) double pipe_smem[];
  SeqWork wk;
  if (!seq_work<false>(p, wk)) return;
  constexpr int NB = D / S;
  constexpr int MP = (MO == 0) ? D : MO;
  constexpr int NST = SeqPipe<D>::NST;
  const int64_t b = wk.b, T = wk.T;
  const bool active = wk.active;
  const int lane = threadIdx.x & 31;
  double* ring = pipe_smem + (size_t)(threadIdx.x >> 5) * SeqPipe<D>::PER_WARP;
  double* tile = ring + NST * SeqPipe<D>::STAGE;
  const int64_t sts = p.sts;

  double Pinf[D][D], lam[NB], Ho[MP][D];
  if (GIVEN == 0) load_mat<D>(p.Pinf + b * p.Pinf_bs, Pinf);
  if (GIVEN != 1) {
#pragma unroll
    for (int i = 0; i < NB; ++i) lam[i] = p.lam[b * p.lam_bs + i];
  }
  if (MO != 0) {
#pragma unroll
    for (int a = 0; a < MP; ++a) {
#pragma unroll
      for (int j = 0; j < D; ++j) Ho[a][j] = p.Hout[a * D + j];
    }
  }
  const int64_t row0 = b * p.sbs;
  const int64_t wrow0 = wk.b0 * p.sbs;
  const double* __restrict__ dtp = p.dt + b * p.dt_bs;
  const double* __restrict__ Ap = (GIVEN == 1) ? p.A + b * p.A_bs : nullptr;
  const double* __restrict__ Qp = (GIVEN == 1) ? p.Q + b * p.Q_bs : nullptr;
  const double* __restrict__ mfp = p.mf + row0 * D;
  const double* __restrict__ Pfp = p.Pf + row0 * D * D;
  const double* __restrict__ mfw = p.mf + wrow0 * D;
  const double* __restrict__ Pfw = p.Pf + wrow0 * D * D;
  constexpr int NP = PackedRow<D>::N;
  const double* __restrict__ pkp = PACK ? p.pk + row0 * NP : nullptr;
  const double* __restrict__ pkw = PACK ? p.pk + wrow0 * NP : nullptr;
  double* __restrict__ msp = p.ms + row0 * MP;
  double* __restrict__ Psp = p.Ps + row0 * MP * MP;
  double* __restrict__ msw = p.ms + wrow0 * MP;
  double* __restrict__ Psw = p.Ps + wrow0 * MP * MP;

  auto emit = [&](int64_t k, const double (&ms)[D], const double (&Ps)[D][D]) {
    if constexpr (MO == 0) {
      warp_store_rows<D>(msw + k * sts * D, ms, tile, lane, wk.nvalid);
      warp_store_rows<D * D>(Psw + k * sts * D * D, flat<D>(Ps), tile, lane, wk.nvalid);
    } else {
      double om[MP], oP[MP][MP], HPs[MP][D];
#pragma unroll
      for (int a = 0; a < MP; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(Ho[a][j], ms[j], acc);
        om[a] = acc;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
#pragma unroll
          for (int c = 0; c < D; ++c) t = fma(Ho[a][c], Ps[c][j], t);
          HPs[a][j] = t;
        }
      }
#pragma unroll
      for (int a = 0; a < MP; ++a) {
#pragma unroll
        for (int c = 0; c < MP; ++c) {
          double t = 0.0;
#pragma unroll
          for (int j = 0; j < D; ++j) t = fma(HPs[a][j], Ho[c][j], t);
          oP[a][c] = t;
        }
      }
      if (active) {
        store_vec<MP>(msp + k * sts * MP, om);
        store_mat<MP>(Psp + k * sts * MP * MP, oP);
      }
    }
  };
  auto stage_of = [&](int64_t k) { return ring + (int)(k % NST) * SeqPipe<D>::STAGE; };
  auto issue = [&](int64_t k) {
    double* st = stage_of(k);
    if constexpr (PACK) {
      warp_rows_async<NP>(pkw + k * sts * NP, st, lane, wk.nvalid);
    } else {
      warp_rows_async<D>(mfw + k * sts * D, st, lane, wk.nvalid);
      warp_rows_async<D * D>(Pfw + k * sts * D * D, st + RowTile<D>::SIZE, lane, wk.nvalid);
    }
    seq_cp_async_commit();
  };
  // this lane's filtered moments of a landed stage
  auto own = [&](const double* st, double (&mf)[D], double (&Pf)[D][D]) {
    if constexpr (PACK) {
      double row[NP];
      tile_own_row<NP>(st, lane, row);
      unpack_row<D>(row, mf, Pf);
    } else {
      tile_own_row<D>(st, lane, mf);
      tile_own_row<D * D>(st + RowTile<D>::SIZE, lane, flat<D>(Pf));
    }
  };
  auto front = [&](int64_t k, double dt, double (&mp)[D], double (&Pp)[D][D], double (&G)[D][D]) {
    double mf[D], Pf[D][D];
    own(stage_of(k), mf, Pf);
    Trans<D, S> A;
    if constexpr (GIVEN == 1) {
      double Q[D][D];
      load_trans_dense<D, S>(Ap + k * D * D, A);
      load_mat<D>(Qp + k * D * D, Q);
      rts_front<D, S>(A, Q, false, mf, Pf, p.jitter, mp, Pp, G);
    } else if constexpr (GIVEN == 2) {
      double Q[D][D];
      iwp_trans<D, S>(lam, dt, A, Q);
      rts_front<D, S>(A, Q, false, mf, Pf, p.jitter, mp, Pp, G);
    } else {
      matern_trans<D, S>(lam, dt, A);
      rts_front<D, S>(A, Pinf, true, mf, Pf, p.jitter, mp, Pp, G);
    }
  };
  auto back = [&](int64_t k, const double (&mp)[D], const double (&Pp)[D][D], const double (&G)[D][D],
                  double (&ms)[D], double (&Ps)[D][D]) {
    double mf[D], Pf[D][D];
    own(stage_of(k), mf, Pf);
    rts_back<D>(mf, Pf, mp, Pp, G, ms, Ps);
  };

  // terminal step: smoothed = filtered
  double ms[D], Ps[D][D];
  if constexpr (PACK) {
    double row[NP];
    load_vec<NP>(pkp + (T - 1) * sts * NP, row);
    unpack_row<D>(row, ms, Ps);
  } else {
    load_vec<D>(mfp + (T - 1) * sts * D, ms);
    load_mat<D>(Pfp + (T - 1) * sts * D * D, Ps);
  }
  emit(T - 1, ms, Ps);
  if (T < 2) return;
  // prologue: front(T - 2)
  double mp[D], Pp[D][D], G[D][D];
  issue(T - 2);
  if (T >= 3) issue(T - 3);
  double dt_n = (T >= 3) ? dtp[T - 3] : 0.0;
  {
    if (T >= 3) asm volatile("cp.async.wait_group 1;" ::: "memory"); else seq_cp_async_wait_all();
    __syncwarp();
    front(T - 2, dtp[T - 2], mp, Pp, G);
  }
  // steady state: front(k) and back(k + 1) share the iteration; rows of step k - 1 are in flight
  for (int64_t k = T - 3; k >= 0; --k) {
    const double dt = dt_n;
    seq_cp_async_wait_all();
    __syncwarp();                                      // stage k landed for every lane; stage k + 2 is free again
    if (k >= 1) {
      issue(k - 1);
      dt_n = dtp[k - 1];
    }
    double mp2[D], Pp2[D][D], G2[D][D];
    front(k, dt, mp2, Pp2, G2);
    back(k + 1, mp, Pp, G, ms, Ps);
    emit(k + 1, ms, Ps);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      mp[i] = mp2[i];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        Pp[i][j] = Pp2[i][j];
        G[i][j] = G2[i][j];
      }
    }
  }
  // epilogue: back(0)
  back(0, mp, Pp, G, ms, Ps);
  emit(0, ms, Ps);
}

// ------------------------------------------------------------------------------------------ launch
static inline int pick_block(int64_t B) {
  // Every thread lives for the whole launch (one series each), so what matters is how evenly the warps fall on
  // the 148 SMs: the SM with the most warps finishes last.  Pick the block size (32 / 64 / 128 threads) with the
  // fewest warps on the fullest SM -- e.g. 32,768 series: 1024 one-warp blocks put 7 warps on the fullest SM,
  // 256 four-warp blocks put 8 there (and leave 40 SMs with 4).  Ties go to the larger block.
  int best = 128;
  int64_t best_warps = -1;
  for (int bs = 128; bs >= 32; bs /= 2) {
    const int64_t blocks = (B + bs - 1) / bs;
    const int64_t warps = ((blocks + 147) / 148) * (bs / 32);
    if (best_warps < 0 || warps < best_warps) { best_warps = warps; best = bs; }
  }
  return best;
}

// ------------------------------------------------------------------------- parallel-in-time summaries
// One thread per (series, chunk): folds the chunk's steps into ONE scan element, in registers (the d <= 4
// counterpart of ps_filter_summary_kernel / ps_smooth_summary_kernel in physs_pscan.cu; same element
// layout in global memory: filter [A | C | J | b | eta], smoother [E | L | g], index b * nchunk + c).
template <int D, int S, int M, bool HID, int GIVEN>
__global__ void __launch_bounds__(128) seq_filter_summary_kernel(const SeqFilterArgs p, double* __restrict__ elems) {
  SeqWork wk;
  if (!seq_work<true>(p, wk)) return;
  constexpr int NB = D / S;
  const int64_t b = wk.b, t0 = wk.t0, T = wk.T;
  const int64_t sts = p.sts;
  double Pinf[D][D], H[M][D], lam[NB];
  if (!GIVEN) {
    load_mat<D>(p.Pinf + b * p.Pinf_bs, Pinf);
#pragma unroll
    for (int i = 0; i < NB; ++i) lam[i] = p.lam[b * p.lam_bs + i];
  }
  if (!HID) {
#pragma unroll
    for (int a = 0; a < M; ++a) {
#pragma unroll
      for (int j = 0; j < D; ++j) H[a][j] = p.H[b * p.H_bs + a * D + j];
    }
  }
  double A[D][D], bv[D], C[D][D], J[D][D], eta[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    bv[i] = 0.0;
    eta[i] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      A[i][j] = (i == j) ? 1.0 : 0.0;
      C[i][j] = 0.0;
      J[i][j] = 0.0;
    }
  }
  const int64_t row0 = b * p.sbs + t0 * sts;
  const double* __restrict__ dtp = p.dt + b * p.dt_bs + t0;
  const double* __restrict__ Yp = p.Y + row0 * M;
  const double* __restrict__ Rp = p.R + b * p.R_bs + t0 * p.R_ts;
  const double* __restrict__ Ap = GIVEN ? p.A + b * p.A_bs + t0 * D * D : nullptr;
  const double* __restrict__ Qp = GIVEN ? p.Q + b * p.Q_bs + t0 * D * D : nullptr;
  double y_n[M], R_n[M][M], dt_n;
  load_vec<M>(Yp, y_n);
  load_mat<M>(Rp, R_n);
  dt_n = dtp[0];
  for (int64_t k = 0; k < T; ++k) {
    double y[M], R[M][M];
    const double dt = dt_n;
#pragma unroll
    for (int a = 0; a < M; ++a) {
      y[a] = y_n[a];
#pragma unroll
      for (int c = 0; c < M; ++c) R[a][c] = R_n[a][c];
    }
    if (k + 1 < T) {
      load_vec<M>(Yp + (k + 1) * sts * M, y_n);
      if (p.R_ts != 0) load_mat<M>(Rp + (k + 1) * p.R_ts, R_n);
      dt_n = dtp[k + 1];
    }
    Trans<D, S> Phi;
    if constexpr (GIVEN) {
      double Q[D][D];
      load_trans_dense<D, S>(Ap + k * D * D, Phi);
      load_mat<D>(Qp + k * D * D, Q);
      kf_predict_givenQ<D, S>(Phi, Q, bv, C);
    } else {
      matern_trans<D, S>(lam, dt, Phi);
      kf_predict_stationary<D, S>(Phi, Pinf, bv, C);
    }
    double Abar[D][D];
    Phi.mulL(A, Abar);
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) A[i][j] = Abar[i][j];
    }
    kf_update_summary<D, M, HID>(bv, C, A, J, eta, H, R, y, p.jitter);
  }
  if (wk.active) {
    double* __restrict__ e = elems + (b * p.nchunk + wk.c) * (3 * D * D + 2 * D);
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        e[i * D + j] = A[i][j];
        e[D * D + i * D + j] = C[i][j];
        e[2 * D * D + i * D + j] = J[i][j];
      }
      e[3 * D * D + i] = bv[i];
      e[3 * D * D + D + i] = eta[i];
    }
  }
}

template <int D, int S, int GIVEN>
__global__ void __launch_bounds__(128) seq_smooth_summary_kernel(const SeqSmoothArgs p, double* __restrict__ elems) {
  SeqWork wk;
  if (!seq_work<true>(p, wk)) return;
  constexpr int NB = D / S;
  const int64_t b = wk.b, t0 = wk.t0, T = wk.T;
  const int64_t sts = p.sts;
  double Pinf[D][D], lam[NB];
  if (!GIVEN) {
    load_mat<D>(p.Pinf + b * p.Pinf_bs, Pinf);
#pragma unroll
    for (int i = 0; i < NB; ++i) lam[i] = p.lam[b * p.lam_bs + i];
  }
  double E[D][D], g[D], L[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    g[i] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      E[i][j] = (i == j) ? 1.0 : 0.0;
      L[i][j] = 0.0;
    }
  }
  const int64_t row0 = b * p.sbs + t0 * sts;
  const double* __restrict__ dtp = p.dt + b * p.dt_bs + t0;
  const double* __restrict__ Ap = GIVEN ? p.A + b * p.A_bs + t0 * D * D : nullptr;
  const double* __restrict__ Qp = GIVEN ? p.Q + b * p.Q_bs + t0 * D * D : nullptr;
  const double* __restrict__ mfp = p.mf + row0 * D;
  const double* __restrict__ Pfp = p.Pf + row0 * D * D;
  double mf_n[D], Pf_n[D][D], dt_n;
  load_vec<D>(mfp + (T - 1) * sts * D, mf_n);
  load_mat<D>(Pfp + (T - 1) * sts * D * D, Pf_n);
  dt_n = dtp[T - 1];
  for (int64_t k = T - 1; k >= 0; --k) {
    double mf[D], Pf[D][D];
    const double dt = dt_n;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      mf[i] = mf_n[i];
#pragma unroll
      for (int j = 0; j < D; ++j) Pf[i][j] = Pf_n[i][j];
    }
    if (k >= 1) {
      load_vec<D>(mfp + (k - 1) * sts * D, mf_n);
      load_mat<D>(Pfp + (k - 1) * sts * D * D, Pf_n);
      dt_n = dtp[k - 1];
    }
    Trans<D, S> Phi;
    if constexpr (GIVEN) {
      double Q[D][D];
      load_trans_dense<D, S>(Ap + k * D * D, Phi);
      load_mat<D>(Qp + k * D * D, Q);
      rts_step<D, S>(Phi, Q, false, mf, Pf, p.jitter, g, L, E);
    } else {
      matern_trans<D, S>(lam, dt, Phi);
      rts_step<D, S>(Phi, Pinf, true, mf, Pf, p.jitter, g, L, E);
    }
  }
  if (wk.active) {
    double* __restrict__ e = elems + (b * p.nchunk + wk.c) * (2 * D * D + D);
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        e[i * D + j] = E[i][j];
        e[D * D + i * D + j] = L[i][j];
      }
      e[2 * D * D + i] = g[i];
    }
  }
}

template <int D, int S, int M, bool HID, int GIVEN>
static int launch_filter_summary(cudaStream_t st, const SeqFilterArgs& a, double* elems) {
  const int64_t n = ((a.B + 31) / 32 * 32) * a.chunk_count;
  const int block = pick_block(n);
  seq_filter_summary_kernel<D, S, M, HID, GIVEN><<<(unsigned)((n + block - 1) / block), block, 0, st>>>(a, elems);
  return cuda_status(cudaGetLastError(), "seq_filter_summary_kernel launch");
}

template <int D, int S, int GIVEN>
static int filter_summary_by_m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid, double* elems) {
  if (hid && m == D) return launch_filter_summary<D, S, D, true, GIVEN>(st, a, elems);
  if (m == 1) return launch_filter_summary<D, S, 1, false, GIVEN>(st, a, elems);
  if (D >= 2 && m == 2) return launch_filter_summary<D, S, (D >= 2 ? 2 : 1), false, GIVEN>(st, a, elems);
  if (D >= 3 && m == 3) return launch_filter_summary<D, S, (D >= 3 ? 3 : 1), false, GIVEN>(st, a, elems);
  if (D >= 4 && m == 4) return launch_filter_summary<D, S, (D >= 4 ? 4 : 1), false, GIVEN>(st, a, elems);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter summary: unsupported observation dim");
}

template <int D, int S, int GIVEN>
static int launch_smooth_summary(cudaStream_t st, const SeqSmoothArgs& a, double* elems) {
  const int64_t n = ((a.B + 31) / 32 * 32) * a.chunk_count;
  const int block = pick_block(n);
  seq_smooth_summary_kernel<D, S, GIVEN><<<(unsigned)((n + block - 1) / block), block, 0, st>>>(a, elems);
  return cuda_status(cudaGetLastError(), "seq_smooth_summary_kernel launch");
}


template <int D, int S, int M, bool HID, int GIVEN>
static int launch_filter(cudaStream_t st, const SeqFilterArgs& a) {
  const int64_t n = ((a.B + 31) / 32 * 32) * (a.nchunk > 0 ? a.chunk_count : 1);
  const int block = pick_block(n) < SeqBlock<D>::THREADS ? pick_block(n) : SeqBlock<D>::THREADS;
  const int64_t grid = (n + block - 1) / block;
  if (a.pk) {
    if constexpr (D <= 4) {
      if (a.nchunk > 0 || a.sbs != 1)
        return set_error(PHYSS_ERR_UNSUPPORTED, "packed hand-over: time-major steps, plain mode");
      seq_filter_kernel<D, S, M, HID, GIVEN, false, true><<<(unsigned)grid, block, 0, st>>>(a);
      return cuda_status(cudaGetLastError(), "seq_filter_kernel (packed) launch");
    } else {
      return set_error(PHYSS_ERR_UNSUPPORTED, "packed hand-over: d <= 4");
    }
  }
  if (a.nchunk > 0)
    seq_filter_kernel<D, S, M, HID, GIVEN, true><<<(unsigned)grid, block, 0, st>>>(a);
  else
    seq_filter_kernel<D, S, M, HID, GIVEN, false><<<(unsigned)grid, block, 0, st>>>(a);
  return cuda_status(cudaGetLastError(), "seq_filter_kernel launch");
}

// series one full wave of a thread-per-series kernel keeps resident (occupancy x threads x SMs)
template <typename K>
static int seq_wave(K kern, int threads, int64_t* out) {
  int blocks = 0, dev = 0, sms = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, threads, 0);
  if (e == cudaSuccess) e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_status(e, "seq kernel occupancy query");
  *out = (int64_t)blocks * threads * sms;
  return PHYSS_OK;
}

template <int D, int S, int MO, int GIVEN>
static int launch_smooth(cudaStream_t st, const SeqSmoothArgs& a) {
  if (a.wave_out) {
    if constexpr (D % 2 == 0) {
      auto kern = seq_smooth_pipe_kernel<D, S, MO, GIVEN>;
      const size_t smem = SeqPipe<D>::smem_bytes(1);
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      int blocks = 0, dev = 0, sms = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, 32, smem);
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      *a.wave_out = (int64_t)blocks * 32 * sms;
      return PHYSS_OK;
    } else {
      return seq_wave(seq_smooth_kernel<D, S, MO, GIVEN, false, true>, 32, a.wave_out);
    }
  }
  const int64_t n = ((a.B + 31) / 32 * 32) * (a.nchunk > 0 ? a.chunk_count : 1);
  const int block = pick_block(n) < SeqBlock<D>::THREADS ? pick_block(n) : SeqBlock<D>::THREADS;
  const int64_t grid = (n + block - 1) / block;
  const bool coal = (a.sbs == 1);
  if constexpr (D <= 4) {
    // (odd d as well: the packed row has an even number of doubles, so the 16-byte ring works where the (m, P) rows
    //  of d = 1, 3 do not)
    if (a.pk && coal && a.nchunk == 0) {
      // packed hand-over from the filter of the same call (same launch shape as the unpacked pipe kernel below)
      auto kern = seq_smooth_pipe_kernel<D, S, MO, GIVEN, true>;
      const size_t smem = SeqPipe<D>::smem_bytes(1);
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
          e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return cuda_status(e, "seq_smooth_pipe_kernel (packed): configuration");
        configured = true;
      }
      kern<<<(unsigned)(n / 32), 32, smem, st>>>(a);
      return cuda_status(cudaGetLastError(), "seq_smooth_pipe_kernel (packed) launch");
    }
  }
  if constexpr (D % 2 == 0) {
    static const bool no_pipe = [] { const char* e = getenv("PHYSS_SEQ_NOPIPE"); return e && e[0] == '1'; }();
    if (!a.pk && coal && a.nchunk == 0 && !no_pipe) {
      // one warp per block: the block scheduler spreads the (few) warps of a batch evenly over the SMs
      auto kern = seq_smooth_pipe_kernel<D, S, MO, GIVEN>;
      const size_t smem = SeqPipe<D>::smem_bytes(1);
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
          e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return cuda_status(e, "seq_smooth_pipe_kernel: configuration");
        configured = true;
      }
      kern<<<(unsigned)(n / 32), 32, smem, st>>>(a);
      return cuda_status(cudaGetLastError(), "seq_smooth_pipe_kernel launch");
    }
  }
  if (a.pk) return set_error(PHYSS_ERR_UNSUPPORTED, "packed hand-over: d <= 4, time-major steps, plain mode");
  if (a.nchunk > 0) {
    if (MO != 0 && a.fixup) return set_error(PHYSS_ERR_BAD_ARG, "smoother fix-up needs full_state output");
    if (coal) seq_smooth_kernel<D, S, MO, GIVEN, true, true><<<(unsigned)grid, block, 0, st>>>(a);
    else seq_smooth_kernel<D, S, MO, GIVEN, true, false><<<(unsigned)grid, block, 0, st>>>(a);
  } else {
    if (coal) seq_smooth_kernel<D, S, MO, GIVEN, false, true><<<(unsigned)grid, block, 0, st>>>(a);
    else seq_smooth_kernel<D, S, MO, GIVEN, false, false><<<(unsigned)grid, block, 0, st>>>(a);
  }
  return cuda_status(cudaGetLastError(), "seq_smooth_kernel launch");
}

// m in 1..D (dense H) or identity H with m == D
template <int D, int S, int GIVEN>
static int filter_by_m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid) {
  if (hid && m == D) return launch_filter<D, S, D, true, GIVEN>(st, a);
  if (m == 1) return launch_filter<D, S, 1, false, GIVEN>(st, a);
  if (D >= 2 && m == 2) return launch_filter<D, S, (D >= 2 ? 2 : 1), false, GIVEN>(st, a);
  if (D >= 3 && m == 3) return launch_filter<D, S, (D >= 3 ? 3 : 1), false, GIVEN>(st, a);
  if (D >= 4 && m == 4) return launch_filter<D, S, (D >= 4 ? 4 : 1), false, GIVEN>(st, a);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter: unsupported observation dim");
}

template <int D, int S, int GIVEN>
static int smooth_by_mo(cudaStream_t st, const SeqSmoothArgs& a, int mo) {
  if (mo == 0) return launch_smooth<D, S, 0, GIVEN>(st, a);
  if (mo == 1) return launch_smooth<D, S, 1, GIVEN>(st, a);
  if (D >= 2 && mo == 2) return launch_smooth<D, S, (D >= 2 ? 2 : 1), GIVEN>(st, a);
  if (D >= 3 && mo == 3) return launch_smooth<D, S, (D >= 3 ? 3 : 1), GIVEN>(st, a);
  if (D >= 4 && mo == 4) return launch_smooth<D, S, (D >= 4 ? 4 : 1), GIVEN>(st, a);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq smoother: unsupported projection dim");
}

}  // namespace physs
