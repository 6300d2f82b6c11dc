// physs_rt_sum_impl.cuh -- chunk-summary kernels of the parallel-in-time scan on register row tiles, as
// templates: instantiated per padded dimension in physs_rt_sum_d8.cu / _d16.cu / _d32.cu.
#pragma once
#include "physs_rt_impl.cuh"

namespace physs {

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

template <int G, int DM, bool GIVEN, int DC = 0, int SC = 0, int MC = 0>
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
  const int d = DC ? DC : L.d, m = MC ? MC : L.m, s = SC ? SC : L.s;   // compile-time shapes as in rt_filter_kernel
  const int nblk = (DC && SC) ? DC / SC : L.nblk;
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
      rt_matern_A<G, DM>(A, s, nblk, lam, dt);
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

template <int G, int DM, bool GIVEN, int DC = 0, int SC = 0>
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
  const int d = DC ? DC : L.d, s = SC ? SC : L.s;
  const int nblk = (DC && SC) ? DC / SC : L.nblk;
  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;
  __syncwarp();
  double* Ls = sm + L.C; double* A = sm + L.A; double* Qm = sm + L.Qm; double* W1 = sm + L.W1; double* W2 = sm + L.W2;
  double* E = sm + L.Acc; double* W3 = sm + L.K; double* W4 = sm + L.J; double* GE = sm + L.GE;
  double* gv = sm + L.vb; double* mpred = sm + L.vbb; double* dm = sm + L.vdm;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;
  for (int i = gl; i < d; i += G) E[i * LD + i] = 1.0;
  if (!GIVEN) {
    g2s<G, DM>(Qm, p.Pinf + bb * p.Pinf_bs, d, d);
    for (int i = gl; i < nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
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
      rt_matern_A<G, DM>(A, s, nblk, lam, dt);
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
int rt_run_filter_summary(cudaStream_t st, const SeqFilterArgs& a, int d, int m, int nblk, bool hid,
                                 int64_t cfirst, int64_t ccount, double* elems) {
  const RtSumLayout L = rt_sum_layout<DM>(d, m, GIVEN ? 0 : nblk, GIVEN, false);
  const size_t per_group = (size_t)L.total * sizeof(double);
  const int64_t ngroups = a.B * ccount;
  auto launch = [&](auto kern) -> int {
    int threads = 0;
    size_t smem = 0;
    const int rc = rt_configure(kern, G, per_group, &threads, &smem, "rt_filter_summary_kernel: configuration");
    if (rc) return rc;
    const int gpb = threads / G;
    const int64_t grid = (ngroups + gpb - 1) / gpb;
    kern<<<(unsigned)grid, threads, smem, st>>>(a, L, hid, cfirst, ccount, elems);
    return cuda_status(cudaGetLastError(), "rt_filter_summary_kernel launch");
  };
  if (!GIVEN && d == DM && L.s == 4) {
    if (m == 1) return launch(rt_filter_summary_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4, GIVEN ? 0 : 1>);
    if (m == d) return launch(rt_filter_summary_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4, GIVEN ? 0 : DM>);
    return launch(rt_filter_summary_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4, 0>);
  }
  return launch(rt_filter_summary_kernel<G, DM, GIVEN>);
}

template <int G, int DM, bool GIVEN>
int rt_run_smooth_summary(cudaStream_t st, const SeqSmoothArgs& a, int d, int nblk, double* elems) {
  const RtSumLayout L = rt_sum_layout<DM>(d, 1, GIVEN ? 0 : nblk, GIVEN, true);
  const size_t per_group = (size_t)L.total * sizeof(double);
  const int64_t ngroups = a.B * a.chunk_count;
  auto launch = [&](auto kern) -> int {
    int threads = 0;
    size_t smem = 0;
    const int rc = rt_configure(kern, G, per_group, &threads, &smem, "rt_smooth_summary_kernel: configuration");
    if (rc) return rc;
    const int gpb = threads / G;
    const int64_t grid = (ngroups + gpb - 1) / gpb;
    kern<<<(unsigned)grid, threads, smem, st>>>(a, L, elems);
    return cuda_status(cudaGetLastError(), "rt_smooth_summary_kernel launch");
  };
  if (!GIVEN && d == DM && L.s == 4) return launch(rt_smooth_summary_kernel<G, DM, GIVEN, GIVEN ? 0 : DM, GIVEN ? 0 : 4>);
  return launch(rt_smooth_summary_kernel<G, DM, GIVEN>);
}


template <int DM>
int rt_filter_summary_dm(cudaStream_t st, bool given, const SeqFilterArgs& a, int d, int m, int nblk, bool hid,
                         int64_t cfirst, int64_t ccount, double* elems);
template <int DM>
int rt_smooth_summary_dm(cudaStream_t st, bool given, const SeqSmoothArgs& a, int d, int nblk, double* elems);

#define PHYSS_RT_SUM_INSTANTIATE(DM_)                                                                     \
  template <>                                                                                             \
  int rt_filter_summary_dm<DM_>(cudaStream_t st, bool given, const SeqFilterArgs& a, int d, int m,        \
                                int nblk, bool hid, int64_t cfirst, int64_t ccount, double* elems) {      \
    return given ? rt_run_filter_summary<DM_, DM_, true>(st, a, d, m, nblk, hid, cfirst, ccount, elems)   \
                 : rt_run_filter_summary<DM_, DM_, false>(st, a, d, m, nblk, hid, cfirst, ccount, elems); \
  }                                                                                                       \
  template <>                                                                                             \
  int rt_smooth_summary_dm<DM_>(cudaStream_t st, bool given, const SeqSmoothArgs& a, int d, int nblk,     \
                                double* elems) {                                                          \
    return given ? rt_run_smooth_summary<DM_, DM_, true>(st, a, d, nblk, elems)                           \
                 : rt_run_smooth_summary<DM_, DM_, false>(st, a, d, nblk, elems);                         \
  }

}  // namespace physs
