// physs_rt_impl.cuh -- kernels of physs_rt.cu (sequential filter / smoother on register row tiles), as
// templates: instantiated per padded dimension in physs_rt_d8.cu / _d16.cu / _d32.cu so that the three
// sizes compile in parallel.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "physs_internal.h"
#include "physs_rt.cuh"

namespace physs {


using namespace rt;

#ifndef PHYSS_RT_DMMA_DEFAULT
#define PHYSS_RT_DMMA_DEFAULT 56
#endif

struct RtLayout {
  int d, m, mo, nblk, s;
  int P, Ac, Pc, Qc, W1, W2, K, S, Sj, H, Ho, Rst[2], AQst[2][2], PfS;
  int vm, vmp, vv, vw, vrd, vy[2], vmf[2], vlam, vdm;
  int total;
};

// Shared-memory slab of one series.  The resident matrices set how many series fit on an SM, which is what
// bounds these latency-bound kernels, so the slab holds only what the step needs:
//   filter  : P, W2 (+ K when m > 1)            -- DISC_MATERN predicts in place (P -= Pinf; W2 = A P; P = W2 A^T + Pinf)
//   smoother: Ps, W1, W2, one Pf staging slot   -- Ps - P_pred is formed in place in Ps; the next Pf is fetched
//                                                  after the last use of this one
//   DISC_MATERN: A_k, Pinf and Q_k = Pinf - A_k Pinf A_k^T block-diagonal, compact [DM][CB];
//   DISC_GIVEN : double-buffered A_k, Q_k
template <int DM>
static RtLayout rt_layout(int d, int m, int mo, int nblk, bool given, bool smoother) {
  RtLayout L{};
  L.d = d; L.m = m; L.mo = mo; L.nblk = nblk; L.s = (nblk > 0) ? d / nblk : d;
  constexpr int LD = Dim<DM>::LD;
  constexpr int MAT = Dim<DM>::MAT;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  L.P = take(MAT); L.W2 = take(MAT);
  if (smoother) {
    L.W1 = take(MAT);                                  // right-hand sides -> gain rows
    L.PfS = take(MAT);
    L.vmf[0] = take(LD); L.vmf[1] = take(LD);
    L.Ho = take((mo > 0 ? mo : 0) * LD);
  } else {
    L.K = (m > 1) ? take(MAT) : 0;                     // gain K [d x m] (the scalar update keeps it in vw)
    L.S = take(m * LD); L.Sj = take(m * LD); L.H = take(m * LD);
    L.Rst[0] = take(m * LD); L.Rst[1] = take(m * LD);
    L.vy[0] = take(LD); L.vy[1] = take(LD);
    L.vv = take(LD); L.vw = take(LD);
  }
  if (given) {
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L.AQst[a][b] = take(MAT);
  } else {
    L.Ac = take(DM * CB); L.Pc = take(DM * CB); L.Qc = take(DM * CB);
  }
  L.vm = take(LD); L.vmp = take(LD); L.vdm = take(LD); L.vrd = take(3 * LD);   // rd + two column buffers (chol)
  L.vlam = take(nblk > 0 ? nblk : 1);
  L.total = rt_slab(off);
  return L;
}

// Launch shape of a lane-group kernel: threads per block (32 / 64 / 128) that put the most series on an SM
// by the occupancy calculator (shared-memory slab per series x series per block, registers), with the
// shared-memory carve-out at its maximum.  Ties go to the smaller block (finer tail).
template <typename K>
static inline int rt_configure(K kern, int G, size_t per_group, int* threads_out, size_t* smem_out, const char* what) {
  size_t cap = 0;
  for (int threads = 32; threads <= 128; threads *= 2) {
    const size_t sm = per_group * (threads / G);
    if (threads >= G && sm <= 200 * 1024 && sm > cap) cap = sm;
  }
  if (cap == 0) return set_error(PHYSS_ERR_UNSUPPORTED, "lane-group kernel: shared memory");
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
  if (e != cudaSuccess) return cuda_status(e, what);
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return cuda_status(e, what);
  int best = 0, best_series = -1;
  for (int threads = 32; threads <= 128; threads *= 2) {
    const int gpb = threads / G;
    const size_t sm = per_group * gpb;
    if (gpb < 1 || sm > 200 * 1024) continue;
    int blocks = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, threads, sm);
    if (e != cudaSuccess) return cuda_status(e, what);
    if (blocks * gpb > best_series) { best_series = blocks * gpb; best = threads; }
  }
  if (best == 0 || best_series <= 0) return set_error(PHYSS_ERR_UNSUPPORTED, "lane-group kernel: no resident block");
  *threads_out = best;
  *smem_out = per_group * (best / G);
  return PHYSS_OK;
}

// PHYSS_RT_VERBOSE=1: print the residency of each launch (blocks per SM, series per wave) to stderr -- used to
// size benchmark batches in whole waves
template <typename K>
static inline void rt_report(const char* name, K kern, int threads, size_t smem, int gpb, int64_t groups) {
  static const bool on = [] { const char* e = getenv("PHYSS_RT_VERBOSE"); return e && e[0] == '1'; }();
  if (!on) return;
  int blocks = 0, dev = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, threads, smem);
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  fprintf(stderr, "[physs rt] %s: %d threads, %zu B smem, %d blocks/SM, %d series/SM, wave = %d series, launch = %lld\n",
          name, threads, smem, blocks, blocks * gpb, blocks * gpb * sms, (long long)groups);
}

// (series, chunk) owned by a group (same convention as physs_grp.cu)
struct RtWork {
  int64_t b, c, v, t0, T;
  bool active, chunked;
};
template <typename Args>
__device__ __forceinline__ RtWork rt_work(const Args& p, int64_t gid) {
  RtWork w;
  w.chunked = p.nchunk > 0;
  const int64_t per = w.chunked ? p.chunk_count : 1;
  const int64_t n = p.B * per;
  w.active = gid < n;
  const int64_t g = w.active ? gid : n - 1;
  w.b = g / per;
  w.c = w.chunked ? p.chunk_first + g % per : 0;
  w.v = w.chunked ? w.b * p.nchunk + w.c : w.b;
  w.t0 = w.chunked ? w.c * p.chunk_len : 0;
  w.T = w.chunked ? ((p.chunk_len < p.T - w.t0) ? p.chunk_len : (p.T - w.t0)) : p.T;
  return w;
}

// max over the G lanes of a group
template <int G>
__device__ __forceinline__ double group_max(double x) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// relative agreement of the shared (m, P) with the stored global one, rows split over the lanes
template <int G, int DM>
__device__ __forceinline__ bool rt_agrees(const double* mv, const double* P, int d, const double* __restrict__ om,
                                          const double* __restrict__ oP, double delta) {
  constexpr int LD = Dim<DM>::LD;
  double dP = 0.0, sP = 0.0, dm = 0.0, sm = 0.0;
  for (int i = lane<G>(); i < d; i += G) {
    const double omi = om[i];
    dm = fmax(dm, fabs(mv[i] - omi));
    sm = fmax(sm, fabs(omi));
    for (int j = 0; j < d; ++j) {
      const double o = oP[i * d + j];
      dP = fmax(dP, fabs(P[i * LD + j] - o));
      sP = fmax(sP, fabs(o));
    }
  }
  dP = group_max<G>(dP); sP = group_max<G>(sP); dm = group_max<G>(dm); sm = group_max<G>(sm);
  return (dP <= delta * sP) && (dm <= delta * sm || dm * dm <= delta * delta * sP);
}

// closed-form Matern transition blocks into the diagonal blocks of A (off-block entries stay zero)
template <int G, int DM>
__device__ __forceinline__ void rt_matern_A(double* __restrict__ A, int s, int nblk,
                                            const double* __restrict__ lam, double dt) {
  constexpr int LD = Dim<DM>::LD;
  for (int b = lane<G>(); b < nblk; b += G) {
    double* blk = A + (b * s) * LD + b * s;
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
        for (int j = 0; j < 2; ++j) blk[i * LD + j] = a[i][j];
    } else if (s == 3) {
      double a[3][3];
      MaternExpm<3>::eval(lam[b], dt, a);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) blk[i * LD + j] = a[i][j];
    } else {
      double a[4][4];
      MaternExpm<4>::eval(lam[b], dt, a);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) blk[i * LD + j] = a[i][j];
    }
  }
}

// the same into compact storage: row i of the block-diagonal A at Ac[i * CB + q]
template <int G, int DM>
__device__ __forceinline__ void rt_matern_Ac(double* __restrict__ Ac, int s, int nblk,
                                             const double* __restrict__ lam, double dt) {
  for (int b = lane<G>(); b < nblk; b += G) {
    double* blk = Ac + (b * s) * CB;
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
        for (int j = 0; j < 2; ++j) blk[i * CB + j] = a[i][j];
    } else if (s == 3) {
      double a[3][3];
      MaternExpm<3>::eval(lam[b], dt, a);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) blk[i * CB + j] = a[i][j];
    } else {
      double a[4][4];
      MaternExpm<4>::eval(lam[b], dt, a);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) blk[i * CB + j] = a[i][j];
    }
  }
}

// in-block entries of the block-diagonal Pinf [d x d] (global, dense) into compact storage
template <int G, int DM>
__device__ __forceinline__ void rt_load_Pc(double* __restrict__ Pc, const double* __restrict__ Pinf, int d, int s) {
  for (int i = lane<G>(); i < d; i += G) {
    const int l0 = (i / s) * s;
    for (int q = 0; q < s; ++q) Pc[i * CB + q] = Pinf[i * d + l0 + q];
  }
}

// ------------------------------------------------------------------------------------------ filter
// DC / SC / MC: compile-time state dim / Matern block size / observation dim (0 = read from the layout).
// The full-size Matern-7/2 shapes (d == DM, s == 4) get their own instantiation: loop bounds, block
// indices and the block-size switch fold to constants.
template <int G, int DM, bool GIVEN, int DC = 0, int SC = 0, int MC = 0>
__global__ void rt_filter_kernel(const SeqFilterArgs p, const RtLayout L, const bool hid) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LD = Dim<DM>::LD;
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const RtWork wk = rt_work(p, (int64_t)blockIdx.x * gpb + g_in_block);
  const bool active = wk.active, chunked = wk.chunked;
  if (p.fixup && p.prev_changed && *p.prev_changed == 0) return;   // the previous fix-up pass was a fixed point
  const int64_t bb = wk.b, vs = wk.v, t0 = wk.t0, T = wk.T;
  const int gl = lane<G>();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = DC ? DC : L.d, m = MC ? MC : L.m, s = SC ? SC : L.s;
  const int nblk = (DC && SC) ? DC / SC : L.nblk;

  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;       // zero padding is an invariant
  __syncwarp();
  double* P = sm + L.P; double* Ac = sm + L.Ac; double* Pc = sm + L.Pc;
  double* W2 = sm + L.W2; double* K = sm + L.K;
  double* S = sm + L.S; double* Sj = sm + L.Sj; double* H = sm + L.H;
  double* mv_ = sm + L.vm; double* mp = sm + L.vmp; double* v = sm + L.vv; double* w = sm + L.vw;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;

  if (chunked && p.from_bnd) {
    g2s<G, DM>(P, p.bnd_P + vs * d * d, d, d);
    for (int i = gl; i < d; i += G) mv_[i] = p.bnd_m[vs * d + i];
  } else {
    g2s<G, DM>(P, p.P0 + bb * p.P0_bs, d, d);
    for (int i = gl; i < d; i += G) mv_[i] = p.m0[bb * p.m0_bs + i];
  }
  if (!GIVEN) {
    rt_load_Pc<G, DM>(Pc, p.Pinf + bb * p.Pinf_bs, d, s);
    for (int i = gl; i < nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
  }
  if (!hid) g2s<G, DM>(H, p.H + bb * p.H_bs, m, d);
  const double* dtp = p.dt + bb * p.dt_bs + t0;
  const int64_t sts = p.sts;
  const int64_t row0 = bb * p.sbs + t0 * sts;
  const double* Yp = p.Y + row0 * m;
  const double* Rp = p.R + bb * p.R_bs + t0 * p.R_ts;
  const double* Ap = GIVEN ? p.A + bb * p.A_bs + t0 * d * d : nullptr;
  const double* Qp = GIVEN ? p.Q + bb * p.Q_bs + t0 * d * d : nullptr;
  double* mfp = p.mf + row0 * d;
  double* Pfp = p.Pf + row0 * d * d;
  double* lkp = p.lml_k ? p.lml_k + row0 : nullptr;
  int streak = 0;
  bool done = false;

  auto stage = [&](int64_t k) {
    const int st = (int)(k & 1);
    for (int a = gl; a < m; a += G) grp::cp_async8(sm + L.vy[st] + a, Yp + k * sts * m + a);
    g2s_async<G, DM>(sm + L.Rst[st], Rp + k * p.R_ts, m, m);
    if (GIVEN) {
      g2s_async<G, DM>(sm + L.AQst[st][0], Ap + k * d * d, d, d);
      g2s_async<G, DM>(sm + L.AQst[st][1], Qp + k * d * d, d, d);
    }
    grp::cp_async_commit();
  };

  LmlAcc acc;
  // speculative chunk mode: start `warm` steps early from (m0, P0), discard those steps.  Groups of one warp
  // may own different chunks, so every group runs the same number of warm-up steps (chunk 0 is never
  // launched together with later chunks in this mode, see pscan_filter_spec)
  const int64_t w0 = (chunked && !p.from_bnd && p.warm > 0) ? ((p.warm < t0) ? p.warm : t0) : 0;
  stage(-w0);
  double dt_n = dtp[-w0];
  for (int64_t k = -w0; k < T; ++k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    grp::cp_async_wait_all();
    __syncwarp();
    if (k + 1 < T) { stage(k + 1); dt_n = dtp[k + 1]; }
    const double* y = sm + L.vy[st];
    const double* R = sm + L.Rst[st];
    // ---- predict
    if (GIVEN) {
      const double* Ak = sm + L.AQst[st][0];
      const double* Qk = sm + L.AQst[st][1];
      mv<G, DM, false>(mp, Ak, mv_, d, d, nullptr, 1.0);
      mm_nn<G, DM, false>(W2, Ak, P, d, d, nullptr, 1.0);                 // A P
      __syncwarp();
      mm_nt<G, DM>(P, W2, Ak, d, d, Qk, 1.0);                              // A P A^T + Q
    } else {
      // P_ = Pinf + A (P - Pinf) A^T: the same as A P A^T + (Pinf - A Pinf A^T) without forming Q_k (measured
      // 15 % faster here than the block-wise Q_k the smoother uses: the scalar-update step is short)
      rt_matern_Ac<G, DM>(Ac, s, nblk, lam, dt);
      add_c<G, DM>(P, Pc, d, s, -1.0);                                      // dP = P - Pinf (in place)
      __syncwarp();
      mv_c<G, DM>(mp, Ac, mv_, d, s);
      mm_cn<G, DM>(W2, Ac, P, d, s, nullptr);                               // A dP
      __syncwarp();
      mm_nc<G, DM>(P, W2, Ac, d, s, Pc);                                    // Pinf + A dP A^T
    }
    __syncwarp();
    double det, mahal;
    int nobs = 0;
    for (int a = 0; a < m; ++a) nobs += (y[a] != y[a]) ? 0 : 1;
    if (m == 1) {
      // ---- scalar update
      const bool obs = nobs == 1;
      double hp = 0.0, mu = 0.0;                                             // (P H^T)[i] for own rows; H mp
      if (hid) {
        mu = mp[0];
      } else {
        for (int l = 0; l < d; ++l) mu = fma(H[l], mp[l], mu);
      }
      double sv = 0.0;                                                       // H P H^T
      for (int i = gl; i < d; i += G) {
        if (hid) {
          hp = P[i * LD];
        } else {
          hp = 0.0;
          for (int l = 0; l < d; ++l) hp = fma(P[i * LD + l], H[l], hp);
        }
        hp = obs ? hp : 0.0;
        w[i] = hp;                                                           // gain column, contiguous
        sv = fma(hid ? (i == 0 ? 1.0 : 0.0) : H[i], hp, sv);
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      const double Sv = sv + R[0];
      const double rSj = fast_rcp(Sv + p.jitter);
      const double vv = obs ? (y[0] - mu) : 0.0;
      for (int i = gl; i < d; i += G) w[i] *= rSj;                           // K = P H^T / (S + jitter)
      __syncwarp();
      for (int i = gl; i < d; i += G) {
        const double ki = w[i];
        mv_[i] = fma(ki, vv, mp[i]);
        const double ks = -ki * Sv;
        double2* __restrict__ prow = reinterpret_cast<double2*>(P + i * LD);
        const double2* __restrict__ k2 = reinterpret_cast<const double2*>(w);
#pragma unroll
        for (int j2 = 0; j2 < DM / 2; ++j2) {                                // P -= K S K^T (K zero-padded)
          double2 pr = prow[j2];
          const double2 kk = k2[j2];
          pr.x = fma(ks, kk.x, pr.x);
          pr.y = fma(ks, kk.y, pr.y);
          prow[j2] = pr;
        }
      }
      const double Sl = obs ? Sv : 1.0;
      det = (Sl > 0.0) ? Sl : nan("");                                        // non-PD S: NaN as in the reference
      mahal = vv * vv * fast_rcp(Sl);
    } else {
      // ---- K rows: PHt = P_ H^T with columns of missing observations zeroed  -> K [d x m]
      if (hid) {
        for (int i = gl; i < d; i += G)
          for (int a = 0; a < m; ++a) K[i * LD + a] = (y[a] != y[a]) ? 0.0 : P[i * LD + a];
      } else {
        mm_nt<G, DM>(K, P, H, d, m, nullptr, 1.0);
        __syncwarp();
        for (int i = gl; i < d; i += G)
          for (int a = 0; a < m; ++a)
            if (y[a] != y[a]) K[i * LD + a] = 0.0;
      }
      for (int a = gl; a < m; a += G) {
        double mu;
        if (hid) {
          mu = mp[a];
        } else {
          mu = 0.0;
          for (int l = 0; l < d; ++l) mu = fma(H[a * LD + l], mp[l], mu);
        }
        const double ya = y[a];
        v[a] = (ya != ya) ? 0.0 : (ya - mu);
        w[a] = v[a];
      }
      __syncwarp();
      // S = M H P_ H^T M + R ; Sj = S + jitter I
      if (!hid) mm_nn<G, DM, false>(S, H, K, m, d, nullptr, 1.0);             // H (P H^T M)
      __syncwarp();
      for (int a = gl; a < m; a += G) {
        const bool oa = !(y[a] != y[a]);
        for (int c = 0; c < m; ++c) {
          const double hph = hid ? K[a * LD + c] : S[a * LD + c];
          const double sv = (oa ? hph : 0.0) + R[a * LD + c];
          S[a * LD + c] = sv;
          Sj[a * LD + c] = sv + (a == c ? p.jitter : 0.0);
        }
      }
      __syncwarp();
      chol<G, DM>(Sj, m, rd);
      chol_solve_t<G, DM>(Sj, m, rd, K, d);                                   // K = P H^T (S + jitter)^-1
      __syncwarp();
      mv<G, DM, false>(mv_, K, v, d, m, mp, 1.0);                             // m = m_ + K v
      mm_nn<G, DM, false>(W2, K, S, d, m, nullptr, 1.0);                      // K S
      __syncwarp();
      mm_nt<G, DM>(P, W2, K, d, d, P, -1.0);                                  // P -= K S K^T
      // lml: un-jittered S with missing rows / cols -> identity
      for (int a = gl; a < m; a += G) {
        for (int c = 0; c < m; ++c) {
          const bool keep = !(y[a] != y[a]) && !(y[c] != y[c]);
          Sj[a * LD + c] = keep ? S[a * LD + c] : (a == c ? 1.0 : 0.0);
        }
      }
      __syncwarp();
      det = chol<G, DM>(Sj, m, rd);
      chol_solve_t<G, DM>(Sj, m, rd, w, 1);
      __syncwarp();
      mahal = 0.0;
      for (int a = 0; a < m; ++a) mahal = fma(v[a], w[a], mahal);
    }
    if (!chunked) acc.add(det, mahal, nobs);
    __syncwarp();
    if (k < 0) continue;                                     // warm-up step: nothing is stored
    // ---- outputs (fix-up: compare with what is stored before overwriting it)
    if (chunked && p.fixup) {          // warp-uniform: the comparison shuffles across the whole warp
      const bool ag = rt_agrees<G, DM>(mv_, P, d, mfp + k * sts * d, Pfp + k * sts * d * d, p.delta);
      if (!done) streak = ag ? streak + 1 : 0;
      if (!done && !ag && active && gl == 0 && p.pass_changed) atomicOr(p.pass_changed, 1);
    }
    __syncwarp();
    if (active && !done) {
      for (int i = gl; i < d; i += G) mfp[k * sts * d + i] = mv_[i];
      s2g<G, DM>(Pfp + k * sts * d * d, P, d, d);
      if (lkp && gl == 0) lkp[k * sts] = lml_term(det, mahal, nobs);
    }
    if (chunked && p.fixup) {
      if (streak >= p.patience) done = true;
      if (__all_sync(0xffffffffu, done || !active)) break;
    }
    __syncwarp();
  }
  if (chunked) {
    if (p.fixup && active && !done && gl == 0) atomicOr(p.unconverged, 1);
  } else if (active && gl == 0) {
    p.lml[bb] = acc.value();
  }
}

// ---------------------------------------------------------------------------------------- smoother
template <int G, int DM, bool GIVEN, int DC = 0, int SC = 0, bool DMMA = false>
__global__ void rt_smooth_kernel(const SeqSmoothArgs p, const RtLayout L) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LD = Dim<DM>::LD;
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const RtWork wk = rt_work(p, (int64_t)blockIdx.x * gpb + g_in_block);
  const bool active = wk.active, chunked = wk.chunked;
  const int64_t bb = wk.b, vs = wk.v, t0 = wk.t0, T = wk.T;
  const int gl = lane<G>();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = DC ? DC : L.d, mo = L.mo, s = SC ? SC : L.s;
  const int nblk = (DC && SC) ? DC / SC : L.nblk;
  const int mp_ = (mo == 0) ? d : mo;

  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;
  __syncwarp();
  double* Ps = sm + L.P; double* Ac = sm + L.Ac; double* Pc = sm + L.Pc; double* Qc = sm + L.Qc;
  double* W1 = sm + L.W1; double* W2 = sm + L.W2;
  double* Pf = sm + L.PfS;
  double* Ho = sm + L.Ho;
  double* ms = sm + L.vm; double* mpred = sm + L.vmp; double* dm = sm + L.vdm;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;

  if (!GIVEN) {
    rt_load_Pc<G, DM>(Pc, p.Pinf + bb * p.Pinf_bs, d, s);
    for (int i = gl; i < nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
  }
  if (mo > 0) g2s<G, DM>(Ho, p.Hout, mo, d);
  const double* dtp = p.dt + bb * p.dt_bs + t0;
  const double* Ap = GIVEN ? p.A + bb * p.A_bs + t0 * d * d : nullptr;
  const double* Qp = GIVEN ? p.Q + bb * p.Q_bs + t0 * d * d : nullptr;
  const int64_t sts = p.sts;
  const int64_t row0 = bb * p.sbs + t0 * sts;
  const double* mfp = p.mf + row0 * d;
  const double* Pfp = p.Pf + row0 * d * d;
  double* msp = p.ms + row0 * mp_;
  double* Psp = p.Ps + row0 * mp_ * mp_;
  int streak = 0;
  bool done = false;

  auto emit = [&](int64_t k) {
    if (mo == 0) {
      if (active && !done) {
        for (int i = gl; i < d; i += G) msp[k * sts * d + i] = ms[i];
        s2g<G, DM>(Psp + k * sts * d * d, Ps, d, d);
      }
    } else {
      mm_nn<G, DM, false>(W1, Ho, Ps, mo, d, nullptr, 1.0);                   // Hout Ps  [mo x d]
      __syncwarp();
      if (active) {
        for (int a = gl; a < mo; a += G) {
          double accm = 0.0;
          for (int l = 0; l < d; ++l) accm = fma(Ho[a * LD + l], ms[l], accm);
          msp[k * sts * mo + a] = accm;
          for (int c = 0; c < mo; ++c) {
            double accv = 0.0;
            for (int l = 0; l < d; ++l) accv = fma(W1[a * LD + l], Ho[c * LD + l], accv);
            Psp[k * sts * mo * mo + a * mo + c] = accv;
          }
        }
      }
    }
    __syncwarp();
  };

  // mf / A_k / Q_k of step k are double-buffered (fetched one step ahead at the top of the previous step);
  // Pf has ONE slot: step k's successor is fetched right after the last use of Pf[k], under emit(k)
  auto stage = [&](int64_t k) {
    const int st = (int)(k & 1);
    for (int i = gl; i < d; i += G) grp::cp_async8(sm + L.vmf[st] + i, mfp + k * sts * d + i);
    if (GIVEN) {
      g2s_async<G, DM>(sm + L.AQst[st][0], Ap + k * d * d, d, d);
      g2s_async<G, DM>(sm + L.AQst[st][1], Qp + k * d * d, d, d);
    }
    grp::cp_async_commit();
  };
  auto stage_Pf = [&](int64_t k) {
    g2s_async<G, DM>(Pf, Pfp + k * sts * d * d, d, d);
    grp::cp_async_commit();
  };

  // speculative chunk mode: start w0 steps past the chunk's end from the filtered state there
  const bool spec = chunked && p.warm > 0;
  const int64_t after = p.T - (t0 + T);
  const int64_t w0 = spec ? ((p.warm < after) ? p.warm : after) : 0;
  const bool carried = chunked && !spec && (wk.c < p.nchunk - 1 || p.carry_last);
  if (carried) {
    g2s<G, DM>(Ps, p.bnd_P + vs * d * d, d, d);
    for (int i = gl; i < d; i += G) ms[i] = p.bnd_m[vs * d + i];
  } else {
    g2s<G, DM>(Ps, Pfp + (T - 1 + w0) * sts * d * d, d, d);
    for (int i = gl; i < d; i += G) ms[i] = mfp[(T - 1 + w0) * sts * d + i];
  }
  __syncwarp();
  int64_t kstart = (w0 > 0) ? T - 2 + w0 : T - 1;
  if (!chunked) {
    emit(T - 1);
    kstart = T - 2;
  }
  double dt_n = 0.0;
  if (kstart >= 0) { stage(kstart); stage_Pf(kstart); dt_n = dtp[kstart]; }
  for (int64_t k = kstart; k >= 0; --k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    grp::cp_async_wait_all();
    __syncwarp();
    if (k >= 1) { stage(k - 1); dt_n = dtp[k - 1]; }
    const double* mf = sm + L.vmf[st];
    // W1 <- (A Pf)^T = Pf A^T (Pf symmetric)  [rows: state j of Pf, the right-hand sides of the gain solve];
    // W2 <- P_pred = A Pf A^T + Q
    if (GIVEN) {
      const double* Ak = sm + L.AQst[st][0];
      const double* Qk = sm + L.AQst[st][1];
      mv<G, DM, false>(mpred, Ak, mf, d, d, nullptr, 1.0);
      mm_nt<G, DM>(W1, Pf, Ak, d, d, nullptr, 1.0);                            // Pf A^T
      __syncwarp();
      mm_nn<G, DM, false>(W2, Ak, W1, d, d, Qk, 1.0);                          // A (Pf A^T) + Q
    } else {
      rt_matern_Ac<G, DM>(Ac, s, nblk, lam, dt);
      __syncwarp();
      q_c<G, DM>(Qc, Ac, Pc, d, s);                                             // Q_k = Pinf - A Pinf A^T
      mv_c<G, DM>(mpred, Ac, mf, d, s);
      mm_nc<G, DM>(W1, Pf, Ac, d, s, nullptr);                                  // Pf A^T
      __syncwarp();
      mm_cn<G, DM>(W2, Ac, W1, d, s, Qc);                                       // A (Pf A^T) + Q_k
    }
    __syncwarp();
    // dP = Ps - Pp (in place in Ps) ; dm = ms - mpred ; Pp += jitter I
    sub_rows_inplace<G, DM>(Ps, W2, d);
    for (int i = gl; i < d; i += G) {
      W2[i * LD + i] += p.jitter;
      dm[i] = ms[i] - mpred[i];
    }
    __syncwarp();
    chol<G, DM>(W2, d, rd);
    chol_solve_t<G, DM>(W2, d, rd, W1, d);              // rows: W1[j][:] = (Pp + jit)^-1 (A Pf)[:, j]  = G[j][:]
    __syncwarp();
    mv<G, DM, false>(ms, W1, dm, d, d, mf, 1.0);        // ms = mf + G dm
    if constexpr (DMMA && DC == DM && (DM == 8 || DM == 16 || DM == 32)) {
      // the two dense products of the step on the FP64 tensor cores (DMMA.8x8x4): per series 2 (DM/8)^2 (DM/4)
      // instructions instead of 2 DM^2 DFMA per lane; the fragments of G serve both as the left operand of
      // G dP and as G^T on the right.  All 32 lanes work on one series at a time: the warp walks its 32 / DM
      // series (their slabs are consecutive in shared memory).
      constexpr int NG = 32 / G;
      double* slab0 = smem + (size_t)(g_in_block - (g_in_block % NG)) * L.total;
      double ag[NG][DM / 8][DM / 4];
#pragma unroll
      for (int sgi = 0; sgi < NG; ++sgi) {
        double* sb = slab0 + (size_t)sgi * L.total;
        dmma_load_a<DM>(sb + L.W1, ag[sgi]);
        dmma_mm_nn<DM>(sb + L.W2, ag[sgi], sb + L.P);            // W2 = G dP
      }
      __syncwarp();
#pragma unroll
      for (int sgi = 0; sgi < NG; ++sgi) {
        double* sb = slab0 + (size_t)sgi * L.total;
        dmma_mm_nt<DM>(sb + L.P, sb + L.W2, ag[sgi], sb + L.PfS);   // Ps = Pf + (G dP) G^T
      }
    } else {
      mm_nn<G, DM, false>(W2, W1, Ps, d, d, nullptr, 1.0);   // G dP
      __syncwarp();
      mm_nt<G, DM>(Ps, W2, W1, d, d, Pf, 1.0);            // Ps = Pf + (G dP) G^T
    }
    __syncwarp();
    if (chunked && p.fixup && mo == 0) {
      const bool ag = rt_agrees<G, DM>(ms, Ps, d, msp + k * sts * d, Psp + k * sts * d * d, p.delta);
      if (!done) streak = ag ? streak + 1 : 0;
    }
    __syncwarp();
    if (k >= 1) stage_Pf(k - 1);                             // Pf[k] is dead from here on
    if (k < T) emit(k);                                      // steps past the chunk's end are warm-up
    if (chunked && p.fixup) {
      if (streak >= p.patience) done = true;
      if (__all_sync(0xffffffffu, done || !active)) break;
    }
  }
  if (chunked && p.fixup && active && !done && gl == 0) atomicOr(p.unconverged, 1);
}

template <int G, int DM, bool GIVEN>
int rt_run_filter(cudaStream_t st, const SeqFilterArgs& a, int d, int m, int nblk, bool hid) {
  const RtLayout L = rt_layout<DM>(d, m, 0, GIVEN ? 0 : nblk, GIVEN, false);
  const size_t per_group = (size_t)L.total * sizeof(double);
  const int64_t ngroups = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  auto launch = [&](auto kern) -> int {
    int threads = 0;
    size_t smem = 0;
    const int rc = rt_configure(kern, G, per_group, &threads, &smem, "rt_filter_kernel: configuration");
    if (rc) return rc;
    const int gpb = threads / G;
    const int64_t grid = (ngroups + gpb - 1) / gpb;
    rt_report("filter", kern, threads, smem, gpb, ngroups);
    kern<<<(unsigned)grid, threads, smem, st>>>(a, L, hid);
    return cuda_status(cudaGetLastError(), "rt_filter_kernel launch");
  };
  if (!GIVEN && d == DM && L.s == 4) {
    if (m == 1) return launch(rt_filter_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4, GIVEN ? 0 : 1>);
    if (m == d) return launch(rt_filter_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4, GIVEN ? 0 : DM>);
    return launch(rt_filter_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4, 0>);
  }
  return launch(rt_filter_kernel<G, DM, GIVEN>);
}

template <int G, int DM, bool GIVEN>
int rt_run_smooth(cudaStream_t st, const SeqSmoothArgs& a, int d, int mo, int nblk) {
  const RtLayout L = rt_layout<DM>(d, 1, mo, GIVEN ? 0 : nblk, GIVEN, true);
  const size_t per_group = (size_t)L.total * sizeof(double);
  const int64_t ngroups = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  auto launch = [&](auto kern) -> int {
    int threads = 0;
    size_t smem = 0;
    const int rc = rt_configure(kern, G, per_group, &threads, &smem, "rt_smooth_kernel: configuration");
    if (rc) return rc;
    const int gpb = threads / G;
    if (a.wave_out) {
      int blocks = 0, dev = 0, sms = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, threads, smem);
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      *a.wave_out = (int64_t)blocks * gpb * sms;
      return PHYSS_OK;
    }
    const int64_t grid = (ngroups + gpb - 1) / gpb;
    rt_report("smoother", kern, threads, smem, gpb, ngroups);
    kern<<<(unsigned)grid, threads, smem, st>>>(a, L);
    return cuda_status(cudaGetLastError(), "rt_smooth_kernel launch");
  };
  if (!GIVEN && d == DM && L.s == 4) {
    if constexpr (!GIVEN) {
      // PHYSS_RT_DMMA: bit mask of the padded dims whose dense products run on the FP64 tensor cores
      // (8 | 16 | 32; default: those the A-B timing of DESIGN.md section 3 found faster)
      static const int dmma = [] { const char* e = getenv("PHYSS_RT_DMMA"); return e ? atoi(e) : PHYSS_RT_DMMA_DEFAULT; }();
      if (dmma & DM) return launch(rt_smooth_kernel<G, DM, GIVEN, DM, 4, true>);
    }
    return launch(rt_smooth_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4>);
  }
  return launch(rt_smooth_kernel<G, DM, GIVEN>);
}


// per-DM entry points (explicit specialisations live in physs_rt_d*.cu)
template <int DM>
int rt_filter_dm(cudaStream_t st, bool given, const SeqFilterArgs& a, int d, int m, int nblk, bool hid);
template <int DM>
int rt_smooth_dm(cudaStream_t st, bool given, const SeqSmoothArgs& a, int d, int mo, int nblk);

#define PHYSS_RT_INSTANTIATE(DM_)                                                                         \
  template <>                                                                                             \
  int rt_filter_dm<DM_>(cudaStream_t st, bool given, const SeqFilterArgs& a, int d, int m, int nblk,      \
                        bool hid) {                                                                       \
    return given ? rt_run_filter<DM_, DM_, true>(st, a, d, m, nblk, hid)                                  \
                 : rt_run_filter<DM_, DM_, false>(st, a, d, m, nblk, hid);                                \
  }                                                                                                       \
  template <>                                                                                             \
  int rt_smooth_dm<DM_>(cudaStream_t st, bool given, const SeqSmoothArgs& a, int d, int mo, int nblk) {   \
    return given ? rt_run_smooth<DM_, DM_, true>(st, a, d, mo, nblk)                                      \
                 : rt_run_smooth<DM_, DM_, false>(st, a, d, mo, nblk);                                    \
  }

}  // namespace physs
