// physs_rt.cu -- sequential Kalman filter / RTS smoother for state dims 5 .. 32: one lane group per series,
// matrices resident in shared memory (zero-padded to DM = 8 / 16 / 32), products accumulated in register
// row tiles (physs_rt.cuh).  Same reference semantics (kalman_filter.py:144-241,439-485;
// rts_smoother.py:48-106,162-192), ABI and chunk / fix-up modes as physs_grp.cu, which remains the
// fallback for d > 32.
//
//   DM = 8 / 16 / 32 : G = DM lanes per series, one matrix row per lane (4 / 2 / 1 series per warp)
#include <stdlib.h>

#include "physs_internal.h"
#include "physs_rt.cuh"

namespace physs {

using namespace rt;

struct RtLayout {
  int d, m, mo, nblk, s;
  int P, A, Qm, W1, W2, W4, K, S, Sj, H, Ho, Rst[2], AQst[2][2], PfS[2];
  int vm, vmp, vv, vw, vrd, vy[2], vmf[2], vlam, vdm;
  int total;
};

template <int DM>
static RtLayout rt_layout(int d, int m, int mo, int nblk, bool given, bool smoother) {
  RtLayout L{};
  L.d = d; L.m = m; L.mo = mo; L.nblk = nblk; L.s = (nblk > 0) ? d / nblk : d;
  constexpr int LD = Dim<DM>::LD;
  constexpr int MAT = Dim<DM>::MAT;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  L.P = take(MAT); L.A = take(MAT); L.Qm = take(MAT); L.W1 = take(MAT); L.W2 = take(MAT);
  if (smoother) {
    L.K = take(MAT); L.W4 = take(MAT);                 // third / fourth work matrices
    L.PfS[0] = take(MAT); L.PfS[1] = take(MAT);
    L.vmf[0] = take(LD); L.vmf[1] = take(LD);
    L.Ho = take((mo > 0 ? mo : 0) * LD);
  } else {
    L.K = take(MAT);                                   // gain K [d x m]
    L.S = take(m * LD); L.Sj = take(m * LD); L.H = take(m * LD);
    L.Rst[0] = take(m * LD); L.Rst[1] = take(m * LD);
    L.vy[0] = take(LD); L.vy[1] = take(LD);
    L.vv = take(LD); L.vw = take(LD);
  }
  if (given) {
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L.AQst[a][b] = take(MAT);
  }
  L.vm = take(LD); L.vmp = take(LD); L.vdm = take(LD); L.vrd = take(3 * LD);   // rd + two column buffers (chol)
  L.vlam = take(nblk > 0 ? nblk : 1);
  L.total = rt_slab(off);
  return L;
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
      MaternExpm<2>::eval(lam[b], dt, a);
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

// ------------------------------------------------------------------------------------------ filter
template <int G, int DM, bool GIVEN>
__global__ void rt_filter_kernel(const SeqFilterArgs p, const RtLayout L, const bool hid) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LD = Dim<DM>::LD;
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const RtWork wk = rt_work(p, (int64_t)blockIdx.x * gpb + g_in_block);
  const bool active = wk.active, chunked = wk.chunked;
  const int64_t bb = wk.b, vs = wk.v, t0 = wk.t0, T = wk.T;
  const int gl = lane<G>();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = L.d, m = L.m, s = L.s;

  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;       // zero padding is an invariant
  __syncwarp();
  double* P = sm + L.P; double* A = sm + L.A; double* Qm = sm + L.Qm;
  double* W1 = sm + L.W1; double* W2 = sm + L.W2; double* K = sm + L.K;
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
    g2s<G, DM>(Qm, p.Pinf + bb * p.Pinf_bs, d, d);
    for (int i = gl; i < L.nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
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
      rt_matern_A<G, DM>(A, s, L.nblk, lam, dt);
      sub_rows<G, DM>(W1, P, Qm, d);                                        // dP = P - Pinf
      __syncwarp();
      mv<G, DM, false>(mp, A, mv_, d, d, nullptr, 1.0, s);
      mm_nn<G, DM, false>(W2, A, W1, d, d, nullptr, 1.0, s);                // A dP
      __syncwarp();
      mm_nt_blk<G, DM>(P, W2, A, d, d, s, Qm, 1.0);                         // Pinf + A dP A^T
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
      det = Sl;
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
template <int G, int DM, bool GIVEN>
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
  const int d = L.d, mo = L.mo, s = L.s;
  const int mp_ = (mo == 0) ? d : mo;

  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;
  __syncwarp();
  double* Ps = sm + L.P; double* A = sm + L.A; double* Qm = sm + L.Qm;
  double* W1 = sm + L.W1; double* W2 = sm + L.W2; double* W3 = sm + L.K; double* W4 = sm + L.W4;
  double* Ho = sm + L.Ho;
  double* ms = sm + L.vm; double* mpred = sm + L.vmp; double* dm = sm + L.vdm;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;

  if (!GIVEN) {
    g2s<G, DM>(Qm, p.Pinf + bb * p.Pinf_bs, d, d);
    for (int i = gl; i < L.nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
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

  auto stage = [&](int64_t k) {
    const int st = (int)(k & 1);
    for (int i = gl; i < d; i += G) grp::cp_async8(sm + L.vmf[st] + i, mfp + k * sts * d + i);
    g2s_async<G, DM>(sm + L.PfS[st], Pfp + k * sts * d * d, d, d);
    if (GIVEN) {
      g2s_async<G, DM>(sm + L.AQst[st][0], Ap + k * d * d, d, d);
      g2s_async<G, DM>(sm + L.AQst[st][1], Qp + k * d * d, d, d);
    }
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
  if (kstart >= 0) { stage(kstart); dt_n = dtp[kstart]; }
  for (int64_t k = kstart; k >= 0; --k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    grp::cp_async_wait_all();
    __syncwarp();
    if (k >= 1) { stage(k - 1); dt_n = dtp[k - 1]; }
    const double* mf = sm + L.vmf[st];
    const double* Pf = sm + L.PfS[st];
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
      rt_matern_A<G, DM>(A, s, L.nblk, lam, dt);
      sub_rows<G, DM>(W3, Pf, Qm, d);                                           // dPf = Pf - Pinf
      __syncwarp();
      mv<G, DM, false>(mpred, A, mf, d, d, nullptr, 1.0, s);
      mm_nt_blk<G, DM>(W1, Pf, A, d, d, s, nullptr, 1.0);                       // Pf A^T
      mm_nt_blk<G, DM>(W4, W3, A, d, d, s, nullptr, 1.0);                       // dPf A^T
      __syncwarp();
      mm_nn<G, DM, false>(W2, A, W4, d, d, Qm, 1.0, s);                         // Pinf + A dPf A^T
    }
    __syncwarp();
    // dP = Ps - Pp -> W3 ; dm = ms - mpred ; Pp += jitter I
    sub_rows<G, DM>(W3, Ps, W2, d);
    for (int i = gl; i < d; i += G) {
      W2[i * LD + i] += p.jitter;
      dm[i] = ms[i] - mpred[i];
    }
    __syncwarp();
    chol<G, DM>(W2, d, rd);
    chol_solve_t<G, DM>(W2, d, rd, W1, d);              // rows: W1[j][:] = (Pp + jit)^-1 (A Pf)[:, j]  = G[j][:]
    __syncwarp();
    mv<G, DM, false>(ms, W1, dm, d, d, mf, 1.0);        // ms = mf + G dm
    mm_nn<G, DM, false>(W2, W1, W3, d, d, nullptr, 1.0);   // G dP
    __syncwarp();
    mm_nt<G, DM>(Ps, W2, W1, d, d, Pf, 1.0);            // Ps = Pf + (G dP) G^T
    __syncwarp();
    if (chunked && p.fixup && mo == 0) {
      const bool ag = rt_agrees<G, DM>(ms, Ps, d, msp + k * sts * d, Psp + k * sts * d * d, p.delta);
      if (!done) streak = ag ? streak + 1 : 0;
    }
    __syncwarp();
    if (k < T) emit(k);                                      // steps past the chunk's end are warm-up
    if (chunked && p.fixup) {
      if (streak >= p.patience) done = true;
      if (__all_sync(0xffffffffu, done || !active)) break;
    }
  }
  if (chunked && p.fixup && active && !done && gl == 0) atomicOr(p.unconverged, 1);
}

// ------------------------------------------------------------------------- parallel-in-time summaries
// Register-tiled counterparts of ps_filter_summary_kernel / ps_smooth_summary_kernel (physs_pscan.cu): one lane
// group folds the steps of one (series, chunk) into ONE scan element; same element layout in global memory.
struct RtSumLayout {
  int d, m, nblk, s;
  int C, A, Qm, W1, W2, Acc, Abar, J, K, HAt, Zt, HA, S, Sj, H, Rst[2], AQst[2][2], PfS[2], GE;
  int vb, vbb, veta, vv, vw, vrd, vy[2], vmf[2], vlam, vdm;
  int total;
};
template <int DM>
static RtSumLayout rt_sum_layout(int d, int m, int nblk, bool given, bool smoother) {
  RtSumLayout L{};
  L.d = d; L.m = m; L.nblk = nblk; L.s = (nblk > 0) ? d / nblk : d;
  constexpr int LD = Dim<DM>::LD;
  constexpr int MAT = Dim<DM>::MAT;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  L.C = take(MAT); L.A = take(MAT); L.Qm = take(MAT); L.W1 = take(MAT); L.W2 = take(MAT);
  L.Acc = take(MAT); L.Abar = take(MAT);
  if (smoother) {
    L.K = take(MAT); L.J = take(MAT);                  // W3 / W4
    L.PfS[0] = take(MAT); L.PfS[1] = take(MAT); L.GE = take(MAT);
    L.vmf[0] = take(LD); L.vmf[1] = take(LD);
  } else {
    L.J = take(MAT); L.K = take(MAT); L.HAt = take(MAT); L.Zt = take(MAT);
    L.HA = take(m * LD); L.S = take(m * LD); L.Sj = take(m * LD); L.H = take(m * LD);
    L.Rst[0] = take(m * LD); L.Rst[1] = take(m * LD);
    L.vy[0] = take(LD); L.vy[1] = take(LD); L.vv = take(LD); L.vw = take(LD); L.veta = take(LD);
  }
  if (given) {
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L.AQst[a][b] = take(MAT);
  }
  L.vb = take(LD); L.vbb = take(LD); L.vdm = take(LD); L.vrd = take(3 * LD);
  L.vlam = take(nblk > 0 ? nblk : 1);
  L.total = rt_slab(off);
  return L;
}

template <int G, int DM, bool GIVEN>
__global__ void rt_filter_summary_kernel(const SeqFilterArgs p, const RtSumLayout L, const bool hid,
                                         const int64_t cfirst, const int64_t ccount, double* __restrict__ elems) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LD = Dim<DM>::LD;
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = p.B * ccount;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t bb = g / ccount, c = cfirst + g % ccount;
  const int64_t t0 = c * p.chunk_len;
  const int64_t T = (p.chunk_len < p.T - t0) ? p.chunk_len : (p.T - t0);
  const int gl = lane<G>();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = L.d, m = L.m, s = L.s;
  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;
  __syncwarp();
  double* C = sm + L.C; double* A = sm + L.A; double* Qm = sm + L.Qm; double* W1 = sm + L.W1; double* W2 = sm + L.W2;
  double* Acc = sm + L.Acc; double* Abar = sm + L.Abar; double* J = sm + L.J; double* K = sm + L.K;
  double* HAt = sm + L.HAt; double* Zt = sm + L.Zt; double* HA = sm + L.HA;
  double* S = sm + L.S; double* Sj = sm + L.Sj; double* H = sm + L.H;
  double* bv = sm + L.vb; double* bbar = sm + L.vbb; double* eta = sm + L.veta;
  double* v = sm + L.vv; double* w = sm + L.vw; double* rd = sm + L.vrd; double* lam = sm + L.vlam;
  for (int i = gl; i < d; i += G) Acc[i * LD + i] = 1.0;     // conditional element of an empty interval
  if (!GIVEN) {
    g2s<G, DM>(Qm, p.Pinf + bb * p.Pinf_bs, d, d);
    for (int i = gl; i < L.nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
  }
  if (!hid) g2s<G, DM>(H, p.H + bb * p.H_bs, m, d);
  const double* dtp = p.dt + bb * p.dt_bs + t0;
  const int64_t sts = p.sts;
  const int64_t row0 = bb * p.sbs + t0 * sts;
  const double* Yp = p.Y + row0 * m;
  const double* Rp = p.R + bb * p.R_bs + t0 * p.R_ts;
  const double* Ap = GIVEN ? p.A + bb * p.A_bs + t0 * d * d : nullptr;
  const double* Qp = GIVEN ? p.Q + bb * p.Q_bs + t0 * d * d : nullptr;
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
  stage(0);
  double dt_n = dtp[0];
  for (int64_t k = 0; k < T; ++k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    grp::cp_async_wait_all();
    __syncwarp();
    if (k + 1 < T) { stage(k + 1); dt_n = dtp[k + 1]; }
    const double* y = sm + L.vy[st];
    const double* R = sm + L.Rst[st];
    // ---- predict (b, C), Abar = Phi Acc
    if (GIVEN) {
      const double* Ak = sm + L.AQst[st][0];
      const double* Qk = sm + L.AQst[st][1];
      mv<G, DM, false>(bbar, Ak, bv, d, d, nullptr, 1.0);
      mm_nn<G, DM, false>(W2, Ak, C, d, d, nullptr, 1.0);
      mm_nn<G, DM, false>(Abar, Ak, Acc, d, d, nullptr, 1.0);
      __syncwarp();
      mm_nt<G, DM>(C, W2, Ak, d, d, Qk, 1.0);
    } else {
      rt_matern_A<G, DM>(A, s, L.nblk, lam, dt);
      for (int i = gl; i < d; i += G) {
#pragma unroll
        for (int j = 0; j < DM; ++j) W1[i * LD + j] = C[i * LD + j] - Qm[i * LD + j];
      }
      __syncwarp();
      mv<G, DM, false>(bbar, A, bv, d, d, nullptr, 1.0, s);
      mm_nn<G, DM, false>(W2, A, W1, d, d, nullptr, 1.0, s);
      mm_nn<G, DM, false>(Abar, A, Acc, d, d, nullptr, 1.0, s);
      __syncwarp();
      mm_nt_blk<G, DM>(C, W2, A, d, d, s, Qm, 1.0);
    }
    __syncwarp();
    // ---- masked update of (b, C) + rank-m updates of (A, J, eta)
    if (hid) {
      for (int i = gl; i < d; i += G)
        for (int a = 0; a < m; ++a) K[i * LD + a] = (y[a] != y[a]) ? 0.0 : C[i * LD + a];
      for (int a = gl; a < m; a += G) {
        const bool miss = y[a] != y[a];
#pragma unroll
        for (int j = 0; j < DM; ++j) HA[a * LD + j] = miss ? 0.0 : Abar[a * LD + j];
      }
    } else {
      mm_nt<G, DM>(K, C, H, d, m, nullptr, 1.0);
      mm_nn<G, DM, false>(HA, H, Abar, m, d, nullptr, 1.0);
      __syncwarp();
      for (int i = gl; i < d; i += G)
        for (int a = 0; a < m; ++a)
          if (y[a] != y[a]) K[i * LD + a] = 0.0;
      for (int a = gl; a < m; a += G) {
        if (y[a] != y[a]) {
#pragma unroll
          for (int j = 0; j < DM; ++j) HA[a * LD + j] = 0.0;
        }
      }
    }
    for (int a = gl; a < m; a += G) {
      double mu;
      if (hid) {
        mu = bbar[a];
      } else {
        mu = 0.0;
        for (int l = 0; l < d; ++l) mu = fma(H[a * LD + l], bbar[l], mu);
      }
      const double ya = y[a];
      v[a] = (ya != ya) ? 0.0 : (ya - mu);
      w[a] = v[a];
    }
    __syncwarp();
    if (!hid) mm_nn<G, DM, false>(S, H, K, m, d, nullptr, 1.0);
    for (int i = gl; i < d; i += G)                      // HAt = HA^T (rows: state j) ; Zt starts as a copy
      for (int a = 0; a < m; ++a) { const double t = HA[a * LD + i]; HAt[i * LD + a] = t; Zt[i * LD + a] = t; }
    __syncwarp();
    for (int a = gl; a < m; a += G) {
      const bool oa = !(y[a] != y[a]);
      for (int cc = 0; cc < m; ++cc) {
        const double hph = hid ? K[a * LD + cc] : S[a * LD + cc];
        const double sv = (oa ? hph : 0.0) + R[a * LD + cc];
        S[a * LD + cc] = sv;
        Sj[a * LD + cc] = sv + (a == cc ? p.jitter : 0.0);
      }
    }
    __syncwarp();
    chol<G, DM>(Sj, m, rd);
    chol_solve_t<G, DM>(Sj, m, rd, K, d);                   // K rows
    chol_solve_t<G, DM>(Sj, m, rd, Zt, d);                  // Zt[j][:] = (S + jit)^-1 HA[:, j]
    chol_solve_t<G, DM>(Sj, m, rd, w, 1);
    __syncwarp();
    mv<G, DM, false>(bv, K, v, d, m, bbar, 1.0);            // b = bbar + K v
    mv<G, DM, false>(eta, HAt, w, d, m, eta, 1.0);          // eta += HA^T w
    mm_nn<G, DM, false>(W2, K, S, d, m, nullptr, 1.0);      // K S
    mm_nn<G, DM, false>(Acc, K, HA, d, m, Abar, -1.0);      // A = Abar - K HA
    mm_nt<G, DM>(J, HAt, Zt, d, d, J, 1.0);                 // J += HA^T Z
    __syncwarp();
    mm_nt<G, DM>(C, W2, K, d, d, C, -1.0);                  // C -= K S K^T
    __syncwarp();
  }
  if (active) {
    double* e = elems + (bb * p.nchunk + c) * (3LL * d * d + 2 * d);
    s2g<G, DM>(e, Acc, d, d);
    s2g<G, DM>(e + d * d, C, d, d);
    s2g<G, DM>(e + 2 * d * d, J, d, d);
    for (int i = gl; i < d; i += G) { e[3 * d * d + i] = bv[i]; e[3 * d * d + d + i] = eta[i]; }
  }
}

template <int G, int DM, bool GIVEN>
__global__ void rt_smooth_summary_kernel(const SeqSmoothArgs p, const RtSumLayout L, double* __restrict__ elems) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LD = Dim<DM>::LD;
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = p.B * p.chunk_count;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t bb = g / p.chunk_count, c = p.chunk_first + g % p.chunk_count;
  const int64_t t0 = c * p.chunk_len;
  const int64_t T = (p.chunk_len < p.T - t0) ? p.chunk_len : (p.T - t0);
  const int gl = lane<G>();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = L.d, s = L.s;
  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;
  __syncwarp();
  double* Ls = sm + L.C; double* A = sm + L.A; double* Qm = sm + L.Qm; double* W1 = sm + L.W1; double* W2 = sm + L.W2;
  double* E = sm + L.Acc; double* W3 = sm + L.K; double* W4 = sm + L.J; double* GE = sm + L.GE;
  double* gv = sm + L.vb; double* mpred = sm + L.vbb; double* dm = sm + L.vdm;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;
  for (int i = gl; i < d; i += G) E[i * LD + i] = 1.0;
  if (!GIVEN) {
    g2s<G, DM>(Qm, p.Pinf + bb * p.Pinf_bs, d, d);
    for (int i = gl; i < L.nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
  }
  const double* dtp = p.dt + bb * p.dt_bs + t0;
  const double* Ap = GIVEN ? p.A + bb * p.A_bs + t0 * d * d : nullptr;
  const double* Qp = GIVEN ? p.Q + bb * p.Q_bs + t0 * d * d : nullptr;
  const int64_t sts = p.sts;
  const int64_t row0 = bb * p.sbs + t0 * sts;
  const double* mfp = p.mf + row0 * d;
  const double* Pfp = p.Pf + row0 * d * d;
  auto stage = [&](int64_t k) {
    const int st = (int)(k & 1);
    for (int i = gl; i < d; i += G) grp::cp_async8(sm + L.vmf[st] + i, mfp + k * sts * d + i);
    g2s_async<G, DM>(sm + L.PfS[st], Pfp + k * sts * d * d, d, d);
    if (GIVEN) {
      g2s_async<G, DM>(sm + L.AQst[st][0], Ap + k * d * d, d, d);
      g2s_async<G, DM>(sm + L.AQst[st][1], Qp + k * d * d, d, d);
    }
    grp::cp_async_commit();
  };
  stage(T - 1);
  double dt_n = dtp[T - 1];
  for (int64_t k = T - 1; k >= 0; --k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    grp::cp_async_wait_all();
    __syncwarp();
    if (k >= 1) { stage(k - 1); dt_n = dtp[k - 1]; }
    const double* mf = sm + L.vmf[st];
    const double* Pf = sm + L.PfS[st];
    if (GIVEN) {
      const double* Ak = sm + L.AQst[st][0];
      const double* Qk = sm + L.AQst[st][1];
      mv<G, DM, false>(mpred, Ak, mf, d, d, nullptr, 1.0);
      mm_nt<G, DM>(W1, Pf, Ak, d, d, nullptr, 1.0);
      __syncwarp();
      mm_nn<G, DM, false>(W2, Ak, W1, d, d, Qk, 1.0);
    } else {
      rt_matern_A<G, DM>(A, s, L.nblk, lam, dt);
      sub_rows<G, DM>(W3, Pf, Qm, d);
      __syncwarp();
      mv<G, DM, false>(mpred, A, mf, d, d, nullptr, 1.0, s);
      mm_nt_blk<G, DM>(W1, Pf, A, d, d, s, nullptr, 1.0);
      mm_nt_blk<G, DM>(W4, W3, A, d, d, s, nullptr, 1.0);
      __syncwarp();
      mm_nn<G, DM, false>(W2, A, W4, d, d, Qm, 1.0, s);
    }
    __syncwarp();
    sub_rows<G, DM>(W3, Ls, W2, d);
    for (int i = gl; i < d; i += G) {
      W2[i * LD + i] += p.jitter;
      dm[i] = gv[i] - mpred[i];
    }
    __syncwarp();
    chol<G, DM>(W2, d, rd);
    chol_solve_t<G, DM>(W2, d, rd, W1, d);                  // W1 rows = G
    __syncwarp();
    mv<G, DM, false>(gv, W1, dm, d, d, mf, 1.0);            // g = mf + G (g - mpred)
    mm_nn<G, DM, false>(W2, W1, W3, d, d, nullptr, 1.0);    // G dL
    mm_nn<G, DM, false>(GE, W1, E, d, d, nullptr, 1.0);     // G E
    __syncwarp();
    mm_nt<G, DM>(Ls, W2, W1, d, d, Pf, 1.0);                // L = Pf + G dL G^T
    for (int i = gl; i < d; i += G) {
#pragma unroll
      for (int j = 0; j < DM; ++j) E[i * LD + j] = GE[i * LD + j];
    }
    __syncwarp();
  }
  if (active) {
    double* e = elems + (bb * p.nchunk + c) * (2LL * d * d + d);
    s2g<G, DM>(e, E, d, d);
    s2g<G, DM>(e + d * d, Ls, d, d);
    for (int i = gl; i < d; i += G) e[2 * d * d + i] = gv[i];
  }
}

template <int G, int DM, bool GIVEN>
static int rt_run_filter_summary(cudaStream_t st, const SeqFilterArgs& a, int d, int m, int nblk, bool hid,
                                 int64_t cfirst, int64_t ccount, double* elems) {
  const RtSumLayout L = rt_sum_layout<DM>(d, m, GIVEN ? 0 : nblk, GIVEN, false);
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 100 * 1024) threads /= 2;
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024) return set_error(PHYSS_ERR_UNSUPPORTED, "rt filter summary: shared memory");
  const int gpb = threads / G;
  const int64_t grid = (a.B * ccount + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(rt_filter_summary_kernel<G, DM, GIVEN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(rt_filter_summary_kernel)");
  rt_filter_summary_kernel<G, DM, GIVEN><<<(unsigned)grid, threads, smem, st>>>(a, L, hid, cfirst, ccount, elems);
  return cuda_status(cudaGetLastError(), "rt_filter_summary_kernel launch");
}

template <int G, int DM, bool GIVEN>
static int rt_run_smooth_summary(cudaStream_t st, const SeqSmoothArgs& a, int d, int nblk, double* elems) {
  const RtSumLayout L = rt_sum_layout<DM>(d, 1, GIVEN ? 0 : nblk, GIVEN, true);
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 100 * 1024) threads /= 2;
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024) return set_error(PHYSS_ERR_UNSUPPORTED, "rt smoother summary: shared memory");
  const int gpb = threads / G;
  const int64_t grid = (a.B * a.chunk_count + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(rt_smooth_summary_kernel<G, DM, GIVEN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(rt_smooth_summary_kernel)");
  rt_smooth_summary_kernel<G, DM, GIVEN><<<(unsigned)grid, threads, smem, st>>>(a, L, elems);
  return cuda_status(cudaGetLastError(), "rt_smooth_summary_kernel launch");
}

// ---------------------------------------------------------------------------------------- dispatch
bool rt_supported(int d, int m) { return d >= 1 && d <= 32 && m >= 1 && m <= d; }

template <int G, int DM, bool GIVEN>
static int rt_run_filter(cudaStream_t st, const SeqFilterArgs& a, int d, int m, int nblk, bool hid) {
  const RtLayout L = rt_layout<DM>(d, m, 0, GIVEN ? 0 : nblk, GIVEN, false);
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 100 * 1024) threads /= 2;   // >= 2 blocks per SM
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024) return set_error(PHYSS_ERR_UNSUPPORTED, "rt filter: shared memory");
  const int gpb = threads / G;
  const int64_t ngroups = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  const int64_t grid = (ngroups + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(rt_filter_kernel<G, DM, GIVEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(rt_filter_kernel)");
  rt_filter_kernel<G, DM, GIVEN><<<(unsigned)grid, threads, smem, st>>>(a, L, hid);
  return cuda_status(cudaGetLastError(), "rt_filter_kernel launch");
}

template <int G, int DM, bool GIVEN>
static int rt_run_smooth(cudaStream_t st, const SeqSmoothArgs& a, int d, int mo, int nblk) {
  const RtLayout L = rt_layout<DM>(d, 1, mo, GIVEN ? 0 : nblk, GIVEN, true);
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 100 * 1024) threads /= 2;
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024) return set_error(PHYSS_ERR_UNSUPPORTED, "rt smoother: shared memory");
  const int gpb = threads / G;
  const int64_t ngroups = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  const int64_t grid = (ngroups + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(rt_smooth_kernel<G, DM, GIVEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(rt_smooth_kernel)");
  rt_smooth_kernel<G, DM, GIVEN><<<(unsigned)grid, threads, smem, st>>>(a, L);
  return cuda_status(cudaGetLastError(), "rt_smooth_kernel launch");
}

// Default: one row per lane (G = DM).  The step is a chain of short dependent phases (Cholesky columns,
// triangular solves), so the per-series critical path, not the instruction count, sets the speed: measured
// 1.4-1.7x faster than two rows per lane (G = DM / 2) at d = 8 / 16 / 32.  PHYSS_RT_NARROW=1 selects the
// two-rows-per-lane mapping for A-B timing.
static bool rt_narrow() {
  static const bool on = [] { const char* e = getenv("PHYSS_RT_NARROW"); return e && e[0] == '1'; }();
  return on;
}

static int rt_check(int d, int disc_mode, int nblk) {
  if (disc_mode == PHYSS_DISC_GIVEN) return PHYSS_OK;
  const int s = (nblk > 0) ? d / nblk : 0;
  if (s < 1 || s > 4 || s * nblk != d)
    return set_error(PHYSS_ERR_UNSUPPORTED, "DISC_MATERN needs equal blocks of size 1..4");
  return PHYSS_OK;
}

int rt_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, const SeqFilterArgs& a) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
#define RUN(G_, DM_) (given ? rt_run_filter<G_, DM_, true>(st, a, d, m, nblk, hid) \
                            : rt_run_filter<G_, DM_, false>(st, a, d, m, nblk, hid))
  if (rt_narrow()) {
    if (d <= 8) return RUN(4, 8);
    if (d <= 16) return RUN(8, 16);
    return RUN(16, 32);
  }
  if (d <= 8) return RUN(8, 8);
  if (d <= 16) return RUN(16, 16);
  return RUN(32, 32);
#undef RUN
}

int rt_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
#define RUN(G_, DM_) (given ? rt_run_smooth<G_, DM_, true>(st, a, d, mo, nblk) \
                            : rt_run_smooth<G_, DM_, false>(st, a, d, mo, nblk))
  if (rt_narrow()) {
    if (d <= 8) return RUN(4, 8);
    if (d <= 16) return RUN(8, 16);
    return RUN(16, 32);
  }
  if (d <= 8) return RUN(8, 8);
  if (d <= 16) return RUN(16, 16);
  return RUN(32, 32);
#undef RUN
}

int rt_filter_summary(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, const SeqFilterArgs& a,
                      int64_t cfirst, int64_t ccount, double* elems) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
#define RUN(G_, DM_) (given ? rt_run_filter_summary<G_, DM_, true>(st, a, d, m, nblk, hid, cfirst, ccount, elems) \
                            : rt_run_filter_summary<G_, DM_, false>(st, a, d, m, nblk, hid, cfirst, ccount, elems))
  if (d <= 8) return RUN(8, 8);
  if (d <= 16) return RUN(16, 16);
  return RUN(32, 32);
#undef RUN
}

int rt_smooth_summary(cudaStream_t st, int d, int disc_mode, int nblk, const SeqSmoothArgs& a, double* elems) {
  int rc = rt_check(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
#define RUN(G_, DM_) (given ? rt_run_smooth_summary<G_, DM_, true>(st, a, d, nblk, elems) \
                            : rt_run_smooth_summary<G_, DM_, false>(st, a, d, nblk, elems))
  if (d <= 8) return RUN(8, 8);
  if (d <= 16) return RUN(16, 16);
  return RUN(32, 32);
#undef RUN
}

}  // namespace physs
