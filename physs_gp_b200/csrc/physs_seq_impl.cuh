// physs_seq_impl.cuh -- sequential Kalman filter / RTS smoother, one thread per series (d <= 4 fast path,
// d = 6, 8 interim), every d x d block in registers, compile-time unrolled.
//
// Reference call sites replaced: kalman_filter.py:439-485 (filter('sequential')) and
// rts_smoother.py:162-192 (smoother('sequential')); see include/physs_b200.h.
//
// Layout: reference order with a leading batch axis, [B][T][d][d] row-major.  One thread walks one
// series in time; per step it streams y/R/dt in and (m, P) out with 16-byte vector accesses where the
// row is 16-byte aligned.  HBM traffic per state-step is the algorithmic 8*(d*d + d + m*m + m + 1) B
// for the filter and 8*2*(d*d + d) B for the smoother.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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

// ------------------------------------------------------------------------------------------ filter
// relative agreement of a freshly computed (m, P) with the stored one (fix-up pass of chunk mode)
template <int D>
__device__ __forceinline__ bool agrees(const double (&m)[D], const double (&P)[D][D],
                                       const double* __restrict__ om, const double* __restrict__ oP,
                                       double delta) {
  double dP = 0.0, sP = 0.0, dm = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    dm = fmax(dm, fabs(m[i] - om[i]));
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double o = oP[i * D + j];
      dP = fmax(dP, fabs(P[i][j] - o));
      sP = fmax(sP, fabs(o));
    }
  }
  return (dP <= delta * sP) && (dm * dm <= delta * delta * sP);
}

template <int D, int S, int M, bool HID, bool GIVEN, bool CHUNK>
__global__ void __launch_bounds__(128) seq_filter_kernel(const SeqFilterArgs p) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.B * (CHUNK ? p.chunk_count : 1)) return;
  constexpr int NB = D / S;
  const int64_t b = CHUNK ? tid / p.chunk_count : tid;
  const int64_t c = CHUNK ? p.chunk_first + tid % p.chunk_count : 0;
  const int64_t v = CHUNK ? b * p.nchunk + c : b;          // boundary-snapshot index
  const int64_t t0 = CHUNK ? c * p.chunk_len : 0;
  const int64_t Tfull = p.T;
  const int64_t T = CHUNK ? ((p.chunk_len < Tfull - t0) ? p.chunk_len : (Tfull - t0)) : Tfull;

  double m[D], P[D][D], Pinf[D][D], H[M][D], lam[NB];
  if (CHUNK && p.fixup) {
    load_vec<D>(p.bnd_m + v * D, m);
    load_mat<D>(p.bnd_P + v * D * D, P);
  } else {
    load_vec<D>(p.m0 + b * p.m0_bs, m);
    load_mat<D>(p.P0 + b * p.P0_bs, P);
  }
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
  const double* __restrict__ dtp = p.dt + b * p.dt_bs + t0;
  const double* __restrict__ Yp = p.Y + (b * Tfull + t0) * M;
  const double* __restrict__ Rp = p.R + b * p.R_bs + t0 * p.R_ts;
  const double* __restrict__ Ap = GIVEN ? p.A + b * p.A_bs + t0 * D * D : nullptr;
  const double* __restrict__ Qp = GIVEN ? p.Q + b * p.Q_bs + t0 * D * D : nullptr;
  double* __restrict__ mfp = p.mf + (b * Tfull + t0) * D;
  double* __restrict__ Pfp = p.Pf + (b * Tfull + t0) * D * D;
  double* __restrict__ lkp = p.lml_k ? p.lml_k + b * Tfull + t0 : nullptr;
  int streak = 0;

  LmlAcc acc;
  // software prefetch of the next step's streamed inputs
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
      load_vec<M>(Yp + (k + 1) * M, y_n);
      if (p.R_ts != 0) load_mat<M>(Rp + (k + 1) * p.R_ts, R_n);
      dt_n = dtp[k + 1];
    }
    Trans<D, S> A;
    if constexpr (GIVEN) {
      double Q[D][D];
      load_trans_dense<D, S>(Ap + k * D * D, A);
      load_mat<D>(Qp + k * D * D, Q);
      kf_predict_givenQ<D, S>(A, Q, m, P);
    } else {
      matern_trans<D, S>(lam, dt, A);
      kf_predict_stationary<D, S>(A, Pinf, m, P);
    }
    double det, mahal;
    int nobs;
    kf_update<D, M, HID>(m, P, H, R, y, p.jitter, det, mahal, nobs);
    if (!CHUNK) acc.add(det, mahal, nobs);
    if (CHUNK && p.fixup) {
      streak = agrees<D>(m, P, mfp + k * D, Pfp + k * D * D, p.delta) ? streak + 1 : 0;
    }
    store_vec<D>(mfp + k * D, m);
    store_mat<D>(Pfp + k * D * D, P);
    if (lkp) lkp[k] = lml_term(det, mahal, nobs);
    if (CHUNK && p.fixup && streak >= p.patience) return;
  }
  if (CHUNK) {
    if (p.fixup) atomicOr(p.unconverged, 1);
  } else {
    p.lml[b] = acc.value();
  }
}

// ---------------------------------------------------------------------------------------- smoother
// MO == 0: full_state (H = I).  MO > 0: project with Hout [MO, D].
template <int D, int S, int MO, bool GIVEN, bool CHUNK>
__global__ void __launch_bounds__(128) seq_smooth_kernel(const SeqSmoothArgs p) {
  static_assert(!CHUNK || MO == 0, "chunk mode carries and compares the full state");
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.B * (CHUNK ? p.chunk_count : 1)) return;
  constexpr int NB = D / S;
  constexpr int MP = (MO == 0) ? D : MO;
  const int64_t b = CHUNK ? tid / p.chunk_count : tid;
  const int64_t c = CHUNK ? p.chunk_first + tid % p.chunk_count : 0;
  const int64_t v = CHUNK ? b * p.nchunk + c : b;
  const int64_t t0 = CHUNK ? c * p.chunk_len : 0;
  const int64_t Tfull = p.T;
  const int64_t T = CHUNK ? ((p.chunk_len < Tfull - t0) ? p.chunk_len : (Tfull - t0)) : Tfull;

  double Pinf[D][D], lam[NB], Ho[MP][D];
  if (!GIVEN) {
    load_mat<D>(p.Pinf + b * p.Pinf_bs, Pinf);
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
  const double* __restrict__ dtp = p.dt + b * p.dt_bs + t0;
  const double* __restrict__ Ap = GIVEN ? p.A + b * p.A_bs + t0 * D * D : nullptr;
  const double* __restrict__ Qp = GIVEN ? p.Q + b * p.Q_bs + t0 * D * D : nullptr;
  const double* __restrict__ mfp = p.mf + (b * Tfull + t0) * D;
  const double* __restrict__ Pfp = p.Pf + (b * Tfull + t0) * D * D;
  double* __restrict__ msp = p.ms + (b * Tfull + t0) * MP;
  double* __restrict__ Psp = p.Ps + (b * Tfull + t0) * MP * MP;

  auto emit = [&](int64_t k, const double (&ms)[D], const double (&Ps)[D][D]) {
    if (MO == 0) {
      store_vec<D>(msp + k * D, ms);
      store_mat<D>(Psp + k * D * D, Ps);
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
      store_vec<MP>(msp + k * MP, om);
      store_mat<MP>(Psp + k * MP * MP, oP);
    }
  };

  double ms[D], Ps[D][D];
  // plain mode: the last step is terminal (smoothed = filtered).  Chunk mode: every step of the chunk
  // is an RTS step from the carried state of the next chunk's first step; the very last chunk carries
  // its own last filtered state across dt = 0, which reproduces the terminal condition.
  const bool carried = CHUNK && c < p.nchunk - 1;
  int64_t kstart;
  if (CHUNK) {
    if (carried) {
      load_vec<D>(p.bnd_m + v * D, ms);
      load_mat<D>(p.bnd_P + v * D * D, Ps);
    } else {
      load_vec<D>(mfp + (T - 1) * D, ms);
      load_mat<D>(Pfp + (T - 1) * D * D, Ps);
    }
    kstart = T - 1;
  } else {
    load_vec<D>(mfp + (T - 1) * D, ms);
    load_mat<D>(Pfp + (T - 1) * D * D, Ps);
    emit(T - 1, ms, Ps);
    kstart = T - 2;
  }
  int streak = 0;

  // prefetch filtered moments one step ahead (addresses do not depend on the state)
  double mf_n[D], Pf_n[D][D], dt_n = 0.0;
  if (kstart >= 0) {
    load_vec<D>(mfp + kstart * D, mf_n);
    load_mat<D>(Pfp + kstart * D * D, Pf_n);
    dt_n = dtp[kstart];
  }
  for (int64_t k = kstart; k >= 0; --k) {
    double mf[D], Pf[D][D];
    const double dt = dt_n;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      mf[i] = mf_n[i];
#pragma unroll
      for (int j = 0; j < D; ++j) Pf[i][j] = Pf_n[i][j];
    }
    if (k >= 1) {
      load_vec<D>(mfp + (k - 1) * D, mf_n);
      load_mat<D>(Pfp + (k - 1) * D * D, Pf_n);
      dt_n = dtp[k - 1];
    }
    Trans<D, S> A;
    if constexpr (GIVEN) {
      double Q[D][D];
      load_trans_dense<D, S>(Ap + k * D * D, A);
      load_mat<D>(Qp + k * D * D, Q);
      rts_step<D, S>(A, Q, false, mf, Pf, p.jitter, ms, Ps);
    } else {
      matern_trans<D, S>(lam, dt, A);
      rts_step<D, S>(A, Pinf, true, mf, Pf, p.jitter, ms, Ps);
    }
    if (CHUNK && p.fixup) {
      streak = agrees<D>(ms, Ps, msp + k * D, Psp + k * D * D, p.delta) ? streak + 1 : 0;
    }
    emit(k, ms, Ps);
    if (CHUNK && p.fixup && streak >= p.patience) return;
  }
  if (CHUNK && p.fixup) atomicOr(p.unconverged, 1);
}

// ------------------------------------------------------------------------------------------ launch
static inline int pick_block(int64_t B) {
  // With few series the limiter is FP64 issue per SM sub-partition: spread warps as thinly as
  // possible (one warp per CTA) until every sub-partition of the 148 SMs has one.
  if (B <= 148LL * 4 * 32) return 32;
  if (B <= 148LL * 8 * 32) return 64;
  return 128;
}

template <int D, int S, int M, bool HID, bool GIVEN>
static int launch_filter(cudaStream_t st, const SeqFilterArgs& a) {
  const int64_t n = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  const int block = pick_block(n);
  const int64_t grid = (n + block - 1) / block;
  if (a.nchunk > 0)
    seq_filter_kernel<D, S, M, HID, GIVEN, true><<<(unsigned)grid, block, 0, st>>>(a);
  else
    seq_filter_kernel<D, S, M, HID, GIVEN, false><<<(unsigned)grid, block, 0, st>>>(a);
  return cuda_status(cudaGetLastError(), "seq_filter_kernel launch");
}

template <int D, int S, int MO, bool GIVEN>
static int launch_smooth(cudaStream_t st, const SeqSmoothArgs& a) {
  const int64_t n = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  const int block = pick_block(n);
  const int64_t grid = (n + block - 1) / block;
  if (a.nchunk > 0) {
    if constexpr (MO == 0) {
      seq_smooth_kernel<D, S, 0, GIVEN, true><<<(unsigned)grid, block, 0, st>>>(a);
    } else {
      return set_error(PHYSS_ERR_BAD_ARG, "chunked smoother needs full_state output");
    }
  } else {
    seq_smooth_kernel<D, S, MO, GIVEN, false><<<(unsigned)grid, block, 0, st>>>(a);
  }
  return cuda_status(cudaGetLastError(), "seq_smooth_kernel launch");
}

// m in 1..D (dense H) or identity H with m == D
template <int D, int S, bool GIVEN>
static int filter_by_m(cudaStream_t st, const SeqFilterArgs& a, int m, bool hid) {
  if (hid && m == D) return launch_filter<D, S, D, true, GIVEN>(st, a);
  if (m == 1) return launch_filter<D, S, 1, false, GIVEN>(st, a);
  if (D >= 2 && m == 2) return launch_filter<D, S, (D >= 2 ? 2 : 1), false, GIVEN>(st, a);
  if (D >= 3 && m == 3) return launch_filter<D, S, (D >= 3 ? 3 : 1), false, GIVEN>(st, a);
  if (D >= 4 && m == 4) return launch_filter<D, S, (D >= 4 ? 4 : 1), false, GIVEN>(st, a);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq filter: unsupported observation dim");
}

template <int D, int S, bool GIVEN>
static int smooth_by_mo(cudaStream_t st, const SeqSmoothArgs& a, int mo) {
  if (mo == 0) return launch_smooth<D, S, 0, GIVEN>(st, a);
  if (mo == 1) return launch_smooth<D, S, 1, GIVEN>(st, a);
  if (D >= 2 && mo == 2) return launch_smooth<D, S, (D >= 2 ? 2 : 1), GIVEN>(st, a);
  if (D >= 3 && mo == 3) return launch_smooth<D, S, (D >= 3 ? 3 : 1), GIVEN>(st, a);
  if (D >= 4 && mo == 4) return launch_smooth<D, S, (D >= 4 ? 4 : 1), GIVEN>(st, a);
  return set_error(PHYSS_ERR_UNSUPPORTED, "seq smoother: unsupported projection dim");
}

}  // namespace physs
