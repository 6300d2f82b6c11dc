// physs_vjp_grp.cu -- reverse pass of the sequential Kalman filter's log marginal likelihood for GENERAL state
// and observation dimensions (4 < d <= 32, any 1 <= m <= d; DISC_GIVEN transitions): SURVEY.md section 8 row f1
// at the shapes the reference's VB_NG_ADAM epochs differentiate through -- full-state sites m = d, d = 6 .. 12
// (stgp/trainers/trainer.py:43,128-136; stgp/trainers/standard.py:58-91).  physs_vjp.cu covers d <= 4, m = 1 in
// registers; this file is the shared-memory lane-group form of the same chain rule (oracle/adjoint.py:_update_vjp,
// pinned by torch autograd in tests/test_oracle_adjoint.py).
//
// Mapping: one group of G lanes (8 / 16 / 32 by d) = one series; all matrices of a step in shared memory
// (physs_warp.cuh primitives).  Walks the steps backwards, rebuilds step k's prediction from the STORED filtered
// state of step k - 1 and pushes (m_bar, P_bar) through update and predict.  Outputs per step gA_k, gQ_k
// (every entry an independent variable, as jax.grad treats them) and optionally gR_k; per series gH, sum_k gR_k,
// gm0, gP0.  The chain from (gA, gQ) to kernel hyper-parameters is T-independent algebra on the host side
// (physs_gp_b200/models.py differentiates the closed-form transitions with torch).
#include <stdint.h>

#include "physs_core.cuh"
#include "physs_internal.h"
#include "physs_warp.cuh"

namespace physs {

using namespace grp;

namespace {

struct VjpLayout {
  int d, m, ld, ldm;
  // d x d
  int A, Pp, Pm, Pbar, Pb2, T1, T2;
  // m x d
  int H, Hm, W, X, Xbar, Wbar, Hb, T5;
  // m x m
  int S, Lj, Lm, Sji, Smi, Sbar, R;
  // d x m
  int Kbar, T3, T4;
  // vectors
  int mp, mm, mbar, mb2, v, a, vbar, y, rd, msk;
  int total;
};

VjpLayout make_vjp_layout(int d, int m) {
  VjpLayout L{};
  L.d = d; L.m = m; L.ld = d | 1; L.ldm = m | 1;
  int o = 0;
  auto take = [&](int n) { const int at = o; o += (n + 1) & ~1; return at; };
  const int dd = d * L.ld, md = m * L.ld, mm = m * L.ldm, dm = d * L.ldm;
  L.A = take(dd); L.Pp = take(dd); L.Pm = take(dd); L.Pbar = take(dd); L.Pb2 = take(dd); L.T1 = take(dd); L.T2 = take(dd);
  L.H = take(md); L.Hm = take(md); L.W = take(md); L.X = take(md); L.Xbar = take(md); L.Wbar = take(md); L.Hb = take(md);
  L.T5 = take(md);
  L.S = take(mm); L.Lj = take(mm); L.Lm = take(mm); L.Sji = take(mm); L.Smi = take(mm); L.Sbar = take(mm); L.R = take(mm);
  L.Kbar = take(dm); L.T3 = take(dm); L.T4 = take(dm);
  L.mp = take(d); L.mm = take(d); L.mbar = take(d); L.mb2 = take(d);
  L.v = take(m); L.a = take(m); L.vbar = take(m); L.y = take(m); L.rd = take(m); L.msk = take(m);
  L.total = o;
  return L;
}

template <int G>
__global__ void grp_vjp_kernel(const SeqFilterArgs p, const VjpOut o, const VjpLayout L, const bool hid) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G, gib = threadIdx.x / G;
  const int64_t b_raw = (int64_t)blockIdx.x * gpb + gib;
  const bool active = b_raw < p.B;                           // idle groups of the last warp shadow the last series
  const int64_t b = active ? b_raw : p.B - 1;                // (all lanes of a warp reach every __syncwarp)
  const int gl = Lanes<G>::gl();
  double* sm = smem + (size_t)gib * L.total;
  const int d = L.d, m = L.m, ld = L.ld, ldm = L.ldm;
  double *A = sm + L.A, *Pp = sm + L.Pp, *Pm = sm + L.Pm, *Pbar = sm + L.Pbar, *Pb2 = sm + L.Pb2, *T1 = sm + L.T1,
         *T2 = sm + L.T2;
  double *H = sm + L.H, *Hm = sm + L.Hm, *W = sm + L.W, *X = sm + L.X, *Xbar = sm + L.Xbar, *Wbar = sm + L.Wbar,
         *Hb = sm + L.Hb, *T5 = sm + L.T5;
  double *S = sm + L.S, *Lj = sm + L.Lj, *Lmk = sm + L.Lm, *Sji = sm + L.Sji, *Smi = sm + L.Smi, *Sbar = sm + L.Sbar,
         *R = sm + L.R;
  double *Kbar = sm + L.Kbar, *T3 = sm + L.T3, *T4 = sm + L.T4;
  double *mp = sm + L.mp, *mm_ = sm + L.mm, *mbar = sm + L.mbar, *mb2 = sm + L.mb2, *v = sm + L.v, *a = sm + L.a,
         *vbar = sm + L.vbar, *y = sm + L.y, *rd = sm + L.rd, *msk = sm + L.msk;
  const double gbar = o.g_lml ? o.g_lml[b] : 1.0;
  const double jit = p.jitter;
  const int64_t sbs = p.sbs, sts = p.sts;

  if (hid) {
    for (int idx = gl; idx < m * d; idx += G) H[(idx / d) * ld + idx % d] = (idx / d == idx % d) ? 1.0 : 0.0;
  } else {
    g2s<G>(H, ld, p.H + b * p.H_bs, m, d);
  }
  for (int idx = gl; idx < d * d; idx += G) Pbar[(idx / d) * ld + idx % d] = 0.0;
  for (int i = gl; i < d; i += G) mbar[i] = 0.0;
  for (int idx = gl; idx < m * d; idx += G) Hb[(idx / d) * ld + idx % d] = 0.0;
  __syncwarp();

  for (int64_t k = p.T - 1; k >= 0; --k) {
    const int64_t row = b * sbs + k * sts;
    // ---- load the step: A_k, Q_k (into Pm), the state before the step, y_k, R_k
    g2s<G>(A, ld, p.A + b * p.A_bs + k * d * d, d, d);
    g2s<G>(Pm, ld, p.Q + b * p.Q_bs + k * d * d, d, d);
    if (k > 0) {
      const int64_t prow = b * sbs + (k - 1) * sts;
      g2s<G>(Pp, ld, p.Pf + prow * d * d, d, d);
      for (int i = gl; i < d; i += G) mp[i] = p.mf[prow * d + i];
    } else {
      g2s<G>(Pp, ld, p.P0 + b * p.P0_bs, d, d);
      for (int i = gl; i < d; i += G) mp[i] = p.m0[b * p.m0_bs + i];
    }
    g2s<G>(R, ldm, p.R + b * p.R_bs + k * p.R_ts, m, m);
    for (int i = gl; i < m; i += G) {
      const double yi = p.Y[row * m + i];
      const bool obs = !(yi != yi);
      msk[i] = obs ? 1.0 : 0.0;
      y[i] = obs ? yi : 0.0;
    }
    __syncwarp();
    // ---- forward quantities of the step
    mv<G, false>(mm_, A, ld, mp, d, d, nullptr, 1.0);                                  // m_ = A m
    mm<G, false, false>(T1, ld, A, ld, Pp, ld, d, d, d, nullptr, 0, 1.0);              // T1 = A P
    mm<G, false, true>(T2, ld, A, ld, Pp, ld, d, d, d, nullptr, 0, 1.0);               // T2 = A P^T
    for (int idx = gl; idx < m * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      Hm[i * ld + j] = msk[i] * H[i * ld + j];                                         // Hm = M H
    }
    __syncwarp();
    mm<G, false, true>(Pm, ld, T1, ld, A, ld, d, d, d, Pm, ld, 1.0);                   // P_ = A P A^T + Q
    __syncwarp();
    mv<G, false>(v, Hm, ld, mm_, m, d, y, -1.0);                                       // v = y0 - Hm m_
    mm<G, false, false>(W, ld, Hm, ld, Pm, ld, m, d, d, nullptr, 0, 1.0);              // W = Hm P_
    __syncwarp();
    mm<G, false, true>(S, ldm, W, ld, Hm, ld, m, d, m, R, ldm, 1.0);                   // S = W Hm^T + R
    __syncwarp();
    for (int idx = gl; idx < m * m; idx += G) {
      const int i = idx / m, j = idx - i * m;
      const double s = S[i * ldm + j];
      Lj[i * ldm + j] = s + (i == j ? jit : 0.0);
      Lmk[i * ldm + j] = (msk[i] != 0.0 && msk[j] != 0.0) ? s : (i == j ? 1.0 : 0.0);
      Sji[i * ldm + j] = (i == j) ? 1.0 : 0.0;
      Smi[i * ldm + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp();
    chol<G>(Lj, ldm, m, rd);
    __syncwarp();
    chol_solve<G>(Lj, ldm, m, rd, Sji, ldm, m);                                        // Sji = (S + jitter I)^-1
    __syncwarp();
    chol<G>(Lmk, ldm, m, rd);
    __syncwarp();
    chol_solve<G>(Lmk, ldm, m, rd, Smi, ldm, m);                                       // Smi = mask_to_identity(S)^-1
    __syncwarp();
    mm<G, false, false>(X, ld, Sji, ldm, W, ld, m, m, d, nullptr, 0, 1.0);             // X = Sji W   (K = X^T)
    mv<G, false>(a, Smi, ldm, v, m, m, nullptr, 1.0);                                  // a = Smi v
    __syncwarp();
    // ---- adjoint of the update
    mm<G, false, true>(T3, ldm, Pbar, ld, X, ld, d, d, m, nullptr, 0, 1.0);            // T3 = Pbar K
    mm<G, true, true>(T4, ldm, Pbar, ld, X, ld, d, d, m, nullptr, 0, 1.0);             // T4 = Pbar^T K
    mm<G, false, false>(T5, ld, X, ld, Pbar, ld, m, d, d, nullptr, 0, 1.0);            // T5 = K^T Pbar
    mv<G, false>(vbar, X, ld, mbar, m, d, nullptr, 1.0);                               // K^T mbar
    __syncwarp();
    for (int idx = gl; idx < d * m; idx += G) {                                        // Kbar = mbar v^T - (T3 S^T + T4 S)
      const int i = idx / m, j = idx - i * m;
      double s = mbar[i] * v[j];
      for (int l = 0; l < m; ++l) s -= T3[i * ldm + l] * S[j * ldm + l] + T4[i * ldm + l] * S[l * ldm + j];
      Kbar[i * ldm + j] = s;
    }
    for (int i = gl; i < m; i += G) vbar[i] -= gbar * a[i];
    __syncwarp();
    for (int idx = gl; idx < m * d; idx += G) {                                        // Xbar = Kbar^T
      const int i = idx / d, j = idx - i * d;
      Xbar[i * ld + j] = Kbar[j * ldm + i];
    }
    __syncwarp();
    mm<G, true, false>(Wbar, ld, Sji, ldm, Xbar, ld, m, m, d, nullptr, 0, 1.0);        // Wbar = Sji^T Xbar
    __syncwarp();
    for (int idx = gl; idx < m * m; idx += G) {
      // Sbar = -K^T Pbar K - gbar/2 (Smi - a a^T) o mask - Wbar X^T
      const int i = idx / m, j = idx - i * m;
      double s = 0.0;
      for (int l = 0; l < d; ++l) s -= (T5[i * ld + l] + Wbar[i * ld + l]) * X[j * ld + l];
      s -= 0.5 * gbar * (Smi[i * ldm + j] - a[i] * a[j]) * msk[i] * msk[j];
      Sbar[i * ldm + j] = s;
    }
    __syncwarp();
    {
      double* gRs = o.gR_step ? o.gR_step + row * m * m : nullptr;
      double* gRt = o.gR_sum ? o.gR_sum + b * m * m : nullptr;
      for (int idx = gl; idx < m * m; idx += G) {
        const double s = Sbar[(idx / m) * ldm + idx % m];
        if (gRs && active) gRs[idx] = s;
        if (gRt && active) gRt[idx] += s;                         // one lane owns an entry for the whole launch: no race
      }
    }
    // Hmbar = (Sbar + Sbar^T) W + Wbar P_^T - vbar m_^T ;  gH += M Hmbar
    for (int idx = gl; idx < m * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      double s = -vbar[i] * mm_[j];
      for (int l = 0; l < m; ++l) s += (Sbar[i * ldm + l] + Sbar[l * ldm + i]) * W[l * ld + j];
      for (int l = 0; l < d; ++l) s += Wbar[i * ld + l] * Pm[j * ld + l];
      Hb[i * ld + j] += msk[i] * s;
    }
    // P_bar = Pbar + Hm^T Sbar Hm + Hm^T Wbar   (T5 reused: Sbar Hm + Wbar)
    __syncwarp();
    mm<G, false, false>(T5, ld, Sbar, ldm, Hm, ld, m, m, d, Wbar, ld, 1.0);
    __syncwarp();
    mm<G, true, false>(Pb2, ld, Hm, ld, T5, ld, d, m, d, Pbar, ld, 1.0);
    mv<G, true>(mb2, Hm, ld, vbar, d, m, mbar, -1.0);                                  // m_bar = mbar - Hm^T vbar
    __syncwarp();
    // ---- adjoint of the predict: gQ = P_bar, gA = m_bar m^T + P_bar (A P^T) + P_bar^T (A P)
    {
      double* gQ = o.gQ + row * d * d;
      double* gA = o.gA + row * d * d;
      for (int idx = gl; idx < d * d; idx += G) {
        const int i = idx / d, j = idx - i * d;
        double s = mb2[i] * mp[j];
        for (int l = 0; l < d; ++l) s += Pb2[i * ld + l] * T2[l * ld + j] + Pb2[l * ld + i] * T1[l * ld + j];
        if (active) {
          gQ[idx] = Pb2[i * ld + j];
          gA[idx] = s;
        }
      }
    }
    mv<G, true>(mbar, A, ld, mb2, d, d, nullptr, 1.0);                                 // mbar = A^T m_bar
    mm<G, true, false>(Pm, ld, A, ld, Pb2, ld, d, d, d, nullptr, 0, 1.0);              // (P_ is dead by now)
    __syncwarp();
    mm<G, false, false>(Pbar, ld, Pm, ld, A, ld, d, d, d, nullptr, 0, 1.0);            // Pbar = A^T P_bar A
    __syncwarp();
  }
  if (!active) return;
  for (int idx = gl; idx < d * d; idx += G) o.gP0[b * d * d + idx] = Pbar[(idx / d) * ld + idx % d];
  for (int i = gl; i < d; i += G) o.gm0[b * d + i] = mbar[i];
  if (o.gH)
    for (int idx = gl; idx < m * d; idx += G) o.gH[b * m * d + idx] = Hb[(idx / d) * ld + idx % d];
}

template <int G>
int run_vjp(cudaStream_t st, const SeqFilterArgs& a, const VjpOut& o, const VjpLayout& L, bool hid) {
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 200 * 1024) threads /= 2;
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024) return set_error(PHYSS_ERR_UNSUPPORTED, "kf vjp: state dimension too large for the shared-memory path");
  const int gpb = threads / G;
  const int64_t grid = (a.B + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(grp_vjp_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(grp_vjp_kernel)");
  grp_vjp_kernel<G><<<(unsigned)grid, threads, smem, st>>>(a, o, L, hid);
  return cuda_status(cudaGetLastError(), "grp_vjp_kernel launch");
}

}  // namespace

bool grp_vjp_supported(int d, int m, int disc_mode) {
  return disc_mode == PHYSS_DISC_GIVEN && d >= 1 && d <= 32 && m >= 1 && m <= d;
}

int grp_kf_vjp(cudaStream_t st, int d, int m, bool h_identity, const SeqFilterArgs& a, const VjpOut& o) {
  if (!o.gA || !o.gQ || !o.gm0 || !o.gP0) return set_error(PHYSS_ERR_BAD_ARG, "kf vjp (general d): gA, gQ, gm0, gP0 are required");
  const VjpLayout L = make_vjp_layout(d, m);
  const int G = d <= 8 ? 8 : (d <= 16 ? 16 : 32);
  if (G == 8) return run_vjp<8>(st, a, o, L, h_identity);
  if (G == 16) return run_vjp<16>(st, a, o, L, h_identity);
  return run_vjp<32>(st, a, o, L, h_identity);
}

}  // namespace physs
