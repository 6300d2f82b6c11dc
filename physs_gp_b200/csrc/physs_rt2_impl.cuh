// physs_rt2_impl.cuh -- two-kernel RTS smoother for full-size Matern-7/2 stacks (d = DM in {8, 16, 32}, blocks of 4).
//
// The RTS step of rts_smoother.py:48-106 splits into a part that needs only the FILTERED moments of its own step
//     A_k, m_pred = A mf, P_pred = A Pf A^T + Q_k, G_k = Pf A^T (P_pred + jitter I)^-1          (Cholesky + substitutions)
// and the two-line recursion through time
//     ms_k = mf_k + G_k (ms_{k+1} - m_pred),     Ps_k = Pf_k + G_k (Ps_{k+1} - P_pred) G_k^T.
// The one-kernel smoother (physs_rt_impl.cuh) walks both in one sequential chain per series, and the batch is capped
// by the output memory, so its few resident warps expose every latency of the first part.  Here
//   K1 rt2_gain_kernel: one lane group per (series, step) of a time chunk -- no dependence between steps, as many
//      groups as the GPU holds -- writes G_k, P_pred,k, m_pred,k to a scratch ring;
//   K2 rt2_back_kernel: one warp per 32 / DM series walks the chunk backwards; per step dP = Ps - P_pred, two DMMA
//      products with the fragments of G_k read straight from the scratch (global -> registers), Ps kept in the
//      accumulator registers between steps.
// Same arithmetic per element as the one-kernel path except for the accumulation order of the DMMA products.
#pragma once
#include "physs_rt_impl.cuh"

namespace physs {

struct Rt2Args {
  SeqSmoothArgs p;
  int64_t k0, k1;                          // RTS steps of this chunk: k in [k0, k1), k1 <= T - 1
  int64_t nj;                              // K1: lane groups per series (each takes every nj-th step)
  double* Gs; double* Pps; double* mps;    // scratch, entry (k - k0) * B + b: [d, d], [d, d], [d]
  double* st_m; double* st_P;              // carried smoothed state between chunks [B, d], [B, d, d]
  int first;                               // K2: start from the terminal condition and emit step T - 1
};

struct Rt2Layout {
  int W1, W2, Pf, Ac, Pc, Qc, vmf, vmp, vrd, vlam, total;
};
template <int DM>
static Rt2Layout rt2_gain_layout() {
  Rt2Layout L{};
  constexpr int LD = Dim<DM>::LD, MAT = Dim<DM>::MAT;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  L.W1 = take(MAT); L.W2 = take(MAT); L.Pf = take(MAT);
  L.Ac = take(DM * CB); L.Pc = take(DM * CB); L.Qc = take(DM * CB);
  L.vmf = take(LD); L.vmp = take(LD); L.vrd = take(3 * LD); L.vlam = take(DM / 4);
  L.total = rt_slab(off);
  return L;
}

// ------------------------------------------------------------------------------------------- K1: gains
template <int G, int DM>
__global__ void rt2_gain_kernel(const Rt2Args a, const Rt2Layout L) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LD = Dim<DM>::LD;
  constexpr int d = DM, s = 4, nblk = DM / 4;
  const SeqSmoothArgs& p = a.p;
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t ngroups = p.B * a.nj;
  const bool active = gid < ngroups;
  const int64_t gq = active ? gid : ngroups - 1;
  const int64_t bb = gq / a.nj, j = gq % a.nj;
  const int gl = lane<G>();
  double* sm = smem + (size_t)g_in_block * L.total;
  for (int idx = gl; idx < L.total; idx += G) sm[idx] = 0.0;
  __syncwarp();
  double* W1 = sm + L.W1; double* W2 = sm + L.W2; double* Pf = sm + L.Pf;
  double* Ac = sm + L.Ac; double* Pc = sm + L.Pc; double* Qc = sm + L.Qc;
  double* mf = sm + L.vmf; double* mpred = sm + L.vmp; double* rd = sm + L.vrd; double* lam = sm + L.vlam;
  rt_load_Pc<G, DM>(Pc, p.Pinf + bb * p.Pinf_bs, d, s);
  for (int i = gl; i < nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
  __syncwarp();
  const double* dtp = p.dt + bb * p.dt_bs;
  const int64_t sts = p.sts;
  const int64_t row0 = bb * p.sbs;
  const double* mfp = p.mf + row0 * d;
  const double* Pfp = p.Pf + row0 * d * d;
  for (int64_t k = a.k1 - 1 - j; k >= a.k0; k -= a.nj) {
    g2s_async<G, DM>(Pf, Pfp + k * sts * d * d, d, d);
    for (int i = gl; i < d; i += G) grp::cp_async8(mf + i, mfp + k * sts * d + i);
    grp::cp_async_commit();
    rt_matern_Ac<G, DM>(Ac, s, nblk, lam, dtp[k]);           // overlaps the copy
    __syncwarp();
    q_c<G, DM>(Qc, Ac, Pc, d, s);                              // Q_k = Pinf - A Pinf A^T
    grp::cp_async_wait_all();
    __syncwarp();
    mv_c<G, DM>(mpred, Ac, mf, d, s);
    mm_nc<G, DM>(W1, Pf, Ac, d, s, nullptr);                   // Pf A^T
    __syncwarp();
    mm_cn<G, DM>(W2, Ac, W1, d, s, Qc);                        // P_pred = A (Pf A^T) + Q_k
    __syncwarp();
    const int64_t e = (k - a.k0) * p.B + bb;
    if (active) {
      s2g<G, DM>(a.Pps + e * d * d, W2, d, d);                 // un-jittered P_pred for Ps - P_pred
      for (int i = gl; i < d; i += G) a.mps[e * d + i] = mpred[i];
    }
    __syncwarp();
    for (int i = gl; i < d; i += G) W2[i * LD + i] += p.jitter;
    __syncwarp();
    chol<G, DM>(W2, d, rd);
    chol_solve_t<G, DM>(W2, d, rd, W1, d);                     // rows of W1: G[j][:]
    __syncwarp();
    if (active) s2g<G, DM>(a.Gs + e * d * d, W1, d, d);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------- K2: recursion
// 16-byte global load that may be only 8-byte aligned in the batch-major / odd-offset case is never needed here:
// every row of a d x d block (d a multiple of 8) starts 16-byte aligned.
template <int DM>
__global__ void __launch_bounds__(32) rt2_back_kernel(const Rt2Args a) {
  extern __shared__ __align__(16) double smem[];
  constexpr int LD = Dim<DM>::LD, MAT = Dim<DM>::MAT;
  constexpr int NG = 32 / DM, MT = DM / 8, KT = DM / 4, d = DM, dd = DM * DM;
  constexpr int SLAB = 2 * MAT + 2 * LD;                       // dP | W2 | dm | ms
  const SeqSmoothArgs& p = a.p;
  const int ln = threadIdx.x & 31, g = ln >> 2, t = ln & 3;
  const int64_t sts = p.sts;
  int64_t bs[NG];
  bool act[NG];
#pragma unroll
  for (int sg = 0; sg < NG; ++sg) {
    const int64_t b = (int64_t)blockIdx.x * NG + sg;
    act[sg] = b < p.B;
    bs[sg] = act[sg] ? b : p.B - 1;
  }
  double acc[NG][MT][MT][2];                                   // Ps in accumulator layout, carried through time
  // ---- initial state
#pragma unroll
  for (int sg = 0; sg < NG; ++sg) {
    double* sl = smem + sg * SLAB;
    double* msv = sl + 2 * MAT + LD;
    const int64_t b = bs[sg];
    if (a.first) {
      const int64_t r = b * p.sbs + (p.T - 1) * sts;
      const double* Pf = p.Pf + r * dd;
      double* Po = p.Ps + r * dd;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
          const int o = (8 * mt + g) * DM + 8 * nt + 2 * t;
          const double2 v = *reinterpret_cast<const double2*>(Pf + o);
          acc[sg][mt][nt][0] = v.x; acc[sg][mt][nt][1] = v.y;
          if (act[sg]) *reinterpret_cast<double2*>(Po + o) = v;
        }
      for (int i = ln; i < d; i += 32) {
        const double v = p.mf[r * d + i];
        msv[i] = v;
        if (act[sg]) p.ms[r * d + i] = v;
      }
    } else {
      const double* Pst = a.st_P + b * dd;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
          const double2 v = *reinterpret_cast<const double2*>(Pst + (8 * mt + g) * DM + 8 * nt + 2 * t);
          acc[sg][mt][nt][0] = v.x; acc[sg][mt][nt][1] = v.y;
        }
      for (int i = ln; i < d; i += 32) msv[i] = a.st_m[b * d + i];
    }
  }
  __syncwarp();
  // ---- recursion over the chunk
  for (int64_t k = a.k1 - 1; k >= a.k0; --k) {
    double ag[NG][MT][KT];
#pragma unroll
    for (int sg = 0; sg < NG; ++sg) {
      double* sl = smem + sg * SLAB;
      double* dP = sl; double* dm = sl + 2 * MAT; double* msv = dm + LD;
      const int64_t e = (k - a.k0) * p.B + bs[sg];
      const double* Gk = a.Gs + e * dd;
      const double* Ppk = a.Pps + e * dd;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) ag[sg][mt][kt] = Gk[(8 * mt + g) * DM + 4 * kt + t];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
          const double2 pp = *reinterpret_cast<const double2*>(Ppk + (8 * mt + g) * DM + 8 * nt + 2 * t);
          *reinterpret_cast<double2*>(dP + (8 * mt + g) * LD + 8 * nt + 2 * t) =
              make_double2(acc[sg][mt][nt][0] - pp.x, acc[sg][mt][nt][1] - pp.y);
        }
      for (int i = ln; i < d; i += 32) dm[i] = msv[i] - a.mps[e * d + i];
    }
    __syncwarp();
#pragma unroll
    for (int sg = 0; sg < NG; ++sg) {
      double* sl = smem + sg * SLAB;
      dmma_mm_nn<DM>(sl + MAT, ag[sg], sl);                    // W2 = G dP
    }
    __syncwarp();
#pragma unroll
    for (int sg = 0; sg < NG; ++sg) {
      double* sl = smem + sg * SLAB;
      const double* W2 = sl + MAT; const double* dm = sl + 2 * MAT; double* msv = sl + 2 * MAT + LD;
      const int64_t r = bs[sg] * p.sbs + k * sts;
      const double* Pf = p.Pf + r * dd;
      // Ps = Pf + W2 G^T, accumulators initialised with Pf straight from global memory
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
          const double2 v = *reinterpret_cast<const double2*>(Pf + (8 * mt + g) * DM + 8 * nt + 2 * t);
          acc[sg][mt][nt][0] = v.x; acc[sg][mt][nt][1] = v.y;
        }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        double aw[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) aw[mt] = W2[(8 * mt + g) * LD + 4 * kt + t];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < MT; ++nt) dmma884(acc[sg][mt][nt][0], acc[sg][mt][nt][1], aw[mt], ag[sg][nt][kt]);
      }
      if (act[sg]) {
        double* Po = p.Ps + r * dd;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < MT; ++nt)
            *reinterpret_cast<double2*>(Po + (8 * mt + g) * DM + 8 * nt + 2 * t) =
                make_double2(acc[sg][mt][nt][0], acc[sg][mt][nt][1]);
      }
      // ms = mf + G dm: row 8 mt + g from the fragments of G held by the four lanes t = 0..3
      double part[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        double sacc = 0.0;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) sacc = fma(ag[sg][mt][kt], dm[4 * kt + t], sacc);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        part[mt] = sacc;
      }
      __syncwarp();                                            // every lane has read dm / msv of this step
      if (t == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int row = 8 * mt + g;
          const double v = p.mf[r * d + row] + part[mt];
          msv[row] = v;
          if (act[sg]) p.ms[r * d + row] = v;
        }
      }
    }
    __syncwarp();
  }
  // ---- carry the state to the next (earlier) chunk
#pragma unroll
  for (int sg = 0; sg < NG; ++sg) {
    if (!act[sg]) continue;
    const double* msv = smem + sg * SLAB + 2 * MAT + LD;
    double* Pst = a.st_P + bs[sg] * dd;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < MT; ++nt)
        *reinterpret_cast<double2*>(Pst + (8 * mt + g) * DM + 8 * nt + 2 * t) =
            make_double2(acc[sg][mt][nt][0], acc[sg][mt][nt][1]);
    for (int i = ln; i < d; i += 32) a.st_m[bs[sg] * d + i] = msv[i];
  }
}

// ------------------------------------------------------------------------------------------------ host
template <int DM>
int64_t rt2_workspace_doubles(int64_t B, int64_t steps) {
  return B * (int64_t)(DM * DM + DM) + steps * B * (int64_t)(2 * DM * DM + DM);
}

template <int DM>
int rt2_smooth_dm(cudaStream_t st, const SeqSmoothArgs& p, double* ws, int64_t ws_doubles) {
  constexpr int G = DM;
  const int64_t B = p.B, T = p.T;
  const int64_t state = B * (int64_t)(DM * DM + DM), per_step = B * (int64_t)(2 * DM * DM + DM);
  int64_t Tc = (ws_doubles - state) / per_step;
  if (Tc > T - 1) Tc = T - 1;
  if (T > 1 && Tc < 8) return set_error(PHYSS_ERR_BAD_ARG, "two-kernel smoother: workspace too small");
  Rt2Args a{};
  a.p = p;
  a.st_m = ws; a.st_P = ws + B * DM;
  double* scratch = ws + state;
  const Rt2Layout L = rt2_gain_layout<DM>();
  const size_t per_group = (size_t)L.total * sizeof(double);
  auto k1kern = rt2_gain_kernel<G, DM>;
  int threads = 0;
  size_t smem1 = 0;
  int rc = rt_configure(k1kern, G, per_group, &threads, &smem1, "rt2_gain_kernel: configuration");
  if (rc) return rc;
  const int gpb = threads / G;
  int blocks_per_sm = 0, dev = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k1kern, threads, smem1);
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t capacity = (int64_t)blocks_per_sm * gpb * sms;
  auto k2kern = rt2_back_kernel<DM>;
  constexpr int NG = 32 / DM;
  const size_t smem2 = (size_t)NG * (2 * Dim<DM>::MAT + 2 * Dim<DM>::LD) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k2kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k2kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return cuda_status(e, "rt2_back_kernel: configuration");
    configured = true;
  }
  const unsigned grid2 = (unsigned)((B + NG - 1) / NG);
  int64_t k1 = T - 1;
  a.first = 1;
  do {
    const int64_t k0 = (k1 - Tc > 0) ? k1 - Tc : 0;
    a.k0 = k0; a.k1 = k1;
    a.Gs = scratch; a.Pps = scratch + (k1 - k0) * B * DM * DM; a.mps = a.Pps + (k1 - k0) * B * DM * DM;
    if (k1 > k0) {
      int64_t nj = (3 * capacity + B - 1) / B;                 // ~3 waves of groups per chunk
      if (nj < 1) nj = 1;
      if (nj > k1 - k0) nj = k1 - k0;
      a.nj = nj;
      const int64_t ngroups = B * nj;
      const int64_t grid1 = (ngroups + gpb - 1) / gpb;
      k1kern<<<(unsigned)grid1, threads, smem1, st>>>(a, L);
      rc = cuda_status(cudaGetLastError(), "rt2_gain_kernel launch");
      if (rc) return rc;
    }
    k2kern<<<grid2, 32, smem2, st>>>(a);
    rc = cuda_status(cudaGetLastError(), "rt2_back_kernel launch");
    if (rc) return rc;
    a.first = 0;
    k1 = k0;
  } while (k1 > 0);
  return PHYSS_OK;
}

}  // namespace physs
