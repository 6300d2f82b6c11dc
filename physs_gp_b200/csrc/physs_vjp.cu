// physs_vjp.cu -- reverse pass of the sequential Kalman filter's log marginal likelihood (SURVEY.md section 8,
// row f1): what `jax.jacrev` through filter('sequential') computes for the hyper-parameter steps of the
// reference (stgp/trainers/trainer.py:43,128-136; stgp/trainers/standard.py:58-91), as ONE kernel.
//
// The forward recursion is kalman_filter.py:144-241 (predict, masked update with the jittered gain solve,
// P - K S K^T, masked lml).  The reverse pass walks the steps backwards, rebuilds step k's prediction from the
// STORED filtered state of step k-1 (the forward kernel's outputs: nothing else is saved) and pushes the
// adjoints (m_bar, P_bar) of the filtered state through the step.  Derivation and pins: oracle/adjoint.py,
// tests/test_oracle_adjoint.py (torch autograd through a transcription of the forward recursion).
//
// Mapping: one thread = one series, d <= 4, scalar observations (m = 1): everything in registers, as the
// forward register kernels.  Per step the thread reads (m, P)[k-1], y_k, R_k, dt_k and, in DISC_GIVEN mode,
// A_k / Q_k, and writes gA_k / gQ_k; in DISC_MATERN mode the chain to (lam, Pinf) is folded on chip
// (dA/dlam from the closed forms evaluated on dual numbers) and nothing per-step is written but gR.
#include <stdint.h>

#include "physs_core.cuh"
#include "physs_internal.h"

namespace physs {

template <int D, int S, bool GIVEN>
__global__ void __launch_bounds__(64) kf_vjp_kernel(const SeqFilterArgs p, const VjpOut o) {
  constexpr int NB = D / S;
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int64_t sbs = p.sbs, sts = p.sts;
  const double gbar = o.g_lml ? o.g_lml[b] : 1.0;
  // The state before step k, (m, P)[k-1], is fetched ONE STEP AHEAD with cp.async into a two-stage shared-memory
  // ring laid out [stage][piece][thread] (conflict-free): with 255 registers per thread only 8 warps fit on an SM,
  // so the DRAM latency of a dependent load per step would bound the kernel (measured 4.5x slower without).
  constexpr int PW = (D % 2 == 0) ? 2 : 1;                       // doubles per copy piece (16 B when rows allow)
  constexpr int NPC = (D * D + D) / PW;                          // pieces per step
  constexpr int TPB = 64;                                        // threads per block (launch)
  __shared__ __align__(16) double ring[2][NPC][TPB][PW];
  // per-thread copies of Pinf and of dA/dlam live in shared memory ([entry][thread], conflict-free): in registers
  // they spilled, re-read from global memory every step they stalled the dependency chain (long scoreboard)
  __shared__ double sPinf[GIVEN ? 1 : D * D][TPB];
  __shared__ double sdA[GIVEN ? 1 : D * S][TPB];
  __shared__ double sgP[GIVEN ? 1 : D * D][TPB];                 // running gradient w.r.t. Pinf
  const int tid = threadIdx.x;
  auto issue = [&](int64_t k) {
    const int st = (int)(k & 1);
    const double* pm;
    const double* pP;
    if (k > 0) {
      const int64_t prow = b * sbs + (k - 1) * sts;
      pm = p.mf + prow * D;
      pP = p.Pf + prow * D * D;
    } else {
      pm = p.m0 + b * p.m0_bs;
      pP = p.P0 + b * p.P0_bs;
    }
#pragma unroll
    for (int e = 0; e < D * D / PW; ++e) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(&ring[st][e][tid][0]);
      if (PW == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(pP + 2 * e) : "memory");
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(pP + e) : "memory");
    }
#pragma unroll
    for (int e = 0; e < D / PW; ++e) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(&ring[st][D * D / PW + e][tid][0]);
      if (PW == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(pm + 2 * e) : "memory");
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(pm + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const double jit = p.jitter;

  double h[D];
#pragma unroll
  for (int i = 0; i < D; ++i) h[i] = p.H ? p.H[b * p.H_bs + i] : (i == 0 ? 1.0 : 0.0);
  double lam[NB];
  if (!GIVEN) {
#pragma unroll
    for (int q = 0; q < NB; ++q) lam[q] = p.lam[b * p.lam_bs + q];
#pragma unroll
    for (int e = 0; e < D * D; ++e) sPinf[e][threadIdx.x] = p.Pinf[b * p.Pinf_bs + e];
  }

  double mbar[D], Pbar[D][D], gH[D], glam[NB];
  double gRsum = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    mbar[i] = 0.0;
    gH[i] = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      Pbar[i][j] = 0.0;
      if (!GIVEN) sgP[i * D + j][threadIdx.x] = 0.0;
    }
  }
#pragma unroll
  for (int q = 0; q < NB; ++q) glam[q] = 0.0;

  issue(p.T - 1);
  double y_n = p.Y[b * sbs + (p.T - 1) * sts];
  double R_n = p.R[b * p.R_bs + (p.T - 1) * p.R_ts];
  double dt_n = p.dt[b * p.dt_bs + p.T - 1];
  for (int64_t k = p.T - 1; k >= 0; --k) {
    const int64_t row = b * sbs + k * sts;
    const double y = y_n, R = R_n, dt = dt_n;
    if (k > 0) {
      issue(k - 1);
      y_n = p.Y[b * sbs + (k - 1) * sts];
      R_n = p.R[b * p.R_bs + (k - 1) * p.R_ts];
      dt_n = p.dt[b * p.dt_bs + k - 1];
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // ---- state before the step (own pieces of the ring: no block-level synchronisation needed)
    double m[D], P[D][D];
    {
      const int st = (int)(k & 1);
#pragma unroll
      for (int e = 0; e < D * D; ++e) P[e / D][e % D] = ring[st][e / PW][tid][e % PW];
#pragma unroll
      for (int e = 0; e < D; ++e) m[e] = ring[st][D * D / PW + e / PW][tid][e % PW];
    }
    // ---- transition (dense D x D; off-block entries are zero in DISC_MATERN mode)
    double A[D][D], Q[D][D];
    if (GIVEN) {
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          A[i][j] = p.A[b * p.A_bs + k * D * D + i * D + j];
          Q[i][j] = p.Q[b * p.Q_bs + k * D * D + i * D + j];
        }
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) A[i][j] = 0.0;
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        Dual a[S][S];                                          // closed form and its d/dlam in one evaluation
        MaternExpm<S>::template evalT<Dual>(Dual{lam[q], 1.0}, dt, a);
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
          for (int j = 0; j < S; ++j) {
            A[q * S + i][q * S + j] = a[i][j].v;
            sdA[GIVEN ? 0 : (q * S + i) * S + j][tid] = a[i][j].d;
          }
      }
    }
    // ---- forward: predict.  DISC_MATERN: P_ = A P A^T + (Pinf - A Pinf A^T) = Pinf + A (P - Pinf) A^T, so that ONE
    // product W = A (P - Pinf) serves the prediction and, below, the adjoint of A through both P_ and Q_k
    double mp[D], AP[D][D], Pp[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < D; ++l) acc = fma(A[i][l], m[l], acc);
      mp[i] = acc;
    }
    if (!GIVEN) {
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) P[i][j] -= sPinf[GIVEN ? 0 : i * D + j][tid];
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < D; ++l) acc = fma(A[i][l], P[l][j], acc);
        AP[i][j] = acc;                                        // A P   (DISC_MATERN: A (P - Pinf))
      }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double acc = GIVEN ? Q[i][j] : sPinf[GIVEN ? 0 : i * D + j][tid];
#pragma unroll
        for (int l = 0; l < D; ++l) acc = fma(AP[i][l], A[j][l], acc);
        Pp[i][j] = acc;
      }
    // ---- reverse: update (scalar observation)
    double mpbar[D], Ppbar[D][D];
    double Rbar = 0.0;
    const bool obs = !(y != y);
    if (obs) {
      double c[D], hmp = 0.0, hc = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(h[j], Pp[j][i], acc);
        c[i] = acc;                                            // W = H P_ (kalman_filter.py: solve(S, M H P_))
        hmp = fma(h[i], mp[i], hmp);
      }
#pragma unroll
      for (int i = 0; i < D; ++i) hc = fma(h[i], c[i], hc);
      const double Sv = hc + R, Sj = Sv + jit;
      const double rSj = 1.0 / Sj, rS = 1.0 / Sv;
      const double v = y - hmp;
      const double alpha = Sv * rSj * rSj;                     // P+ = P_ - K S K^T = P_ - alpha W^T W
      // adjoints
      double Kdotm = 0.0, Kbar_c = 0.0, cPc = 0.0;
      double Pc[D], Ptc[D];                                    // Pbar c, Pbar^T c
#pragma unroll
      for (int i = 0; i < D; ++i) {
        Kdotm = fma(c[i] * rSj, mbar[i], Kdotm);
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          a1 = fma(Pbar[i][j], c[j], a1);
          a2 = fma(Pbar[j][i], c[j], a2);
        }
        Pc[i] = a1;
        Ptc[i] = a2;
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        cPc = fma(c[i], Pc[i], cPc);
        Kbar_c = fma(mbar[i] * v, c[i], Kbar_c);               // (K_bar from the mean) . c
      }
      const double vbar = Kdotm - gbar * v * rS;
      const double abar = -cPc;                                 // alpha_bar
      double Sbar = abar * (rSj * rSj - 2.0 * Sv * rSj * rSj * rSj) - Kbar_c * rSj * rSj
                    + gbar * (-0.5) * (rS - v * v * rS * rS);
      Rbar = Sbar;
      double cbar[D];
#pragma unroll
      for (int i = 0; i < D; ++i) cbar[i] = -alpha * (Pc[i] + Ptc[i]) + mbar[i] * v * rSj + Sbar * h[i];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc = fma(Pp[i][j], cbar[j], acc);      // P_ W_bar^T
        gH[i] += Sbar * c[i] + acc - vbar * mp[i];
        mpbar[i] = mbar[i] - vbar * h[i];
#pragma unroll
        for (int j = 0; j < D; ++j) Ppbar[i][j] = fma(h[i], cbar[j], Pbar[i][j]);   // W = H P_
      }
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        mpbar[i] = mbar[i];
#pragma unroll
        for (int j = 0; j < D; ++j) Ppbar[i][j] = Pbar[i][j];
      }
    }
    if (o.gR_step) o.gR_step[row] = Rbar;
    gRsum += Rbar;
    // ---- reverse: predict.  Abar = mp_bar m^T + (Pp_bar + Pp_bar^T) A P ; m_bar = A^T mp_bar ; P_bar = A^T Pp_bar A
    double Abar[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double acc = mpbar[i] * m[j];
#pragma unroll
        for (int l = 0; l < D; ++l) acc = fma(Ppbar[i][l] + Ppbar[l][i], AP[l][j], acc);
        Abar[i][j] = acc;
      }
    double T1[D][D];                                           // Pp_bar A
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < D; ++l) acc = fma(Ppbar[i][l], A[l][j], acc);
        T1[i][j] = acc;
      }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < D; ++l) acc = fma(A[l][i], mpbar[l], acc);
      mbar[i] = acc;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double a2 = 0.0;
#pragma unroll
        for (int l = 0; l < D; ++l) a2 = fma(A[l][i], T1[l][j], a2);
        Pbar[i][j] = a2;                                       // A^T Pp_bar A
      }
    }
    if (GIVEN) {
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          o.gA[row * D * D + i * D + j] = Abar[i][j];
          o.gQ[row * D * D + i * D + j] = Ppbar[i][j];
        }
    } else {
      // Q_k = Pinf - A Pinf A^T :  Pinf_bar += Q_bar - A^T Q_bar A  (Q_bar = Pp_bar; Pbar already holds A^T Pp_bar A);
      // the -(Q_bar + Q_bar^T) A Pinf part of Abar is already inside Abar through W = A (P - Pinf)
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) sgP[GIVEN ? 0 : i * D + j][tid] += Ppbar[i][j] - Pbar[i][j];
      // chain to lam through the closed forms on dual numbers
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
          for (int j = 0; j < S; ++j) acc = fma(Abar[q * S + i][q * S + j], sdA[GIVEN ? 0 : (q * S + i) * S + j][tid], acc);
        glam[q] += acc;
      }
    }
  }
  // ---- results
#pragma unroll
  for (int i = 0; i < D; ++i) {
    if (o.gH) o.gH[b * D + i] = gH[i];
    if (o.gm0) o.gm0[b * D + i] = mbar[i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (o.gP0) o.gP0[b * D * D + i * D + j] = Pbar[i][j];
      if (!GIVEN && o.gPinf) o.gPinf[b * D * D + i * D + j] = sgP[GIVEN ? 0 : i * D + j][tid];
    }
  }
  if (!GIVEN && o.glam) {
#pragma unroll
    for (int q = 0; q < NB; ++q) o.glam[b * NB + q] = glam[q];
  }
  if (o.gR_sum) o.gR_sum[b] = gRsum;
}

template <int D, int S, bool GIVEN>
static int vjp_launch(cudaStream_t st, const SeqFilterArgs& a, const VjpOut& o) {
  const int threads = 64;                                        // = TPB of the kernel
  const int64_t grid = (a.B + threads - 1) / threads;
  kf_vjp_kernel<D, S, GIVEN><<<(unsigned)grid, threads, 0, st>>>(a, o);
  return cuda_status(cudaGetLastError(), "kf_vjp_kernel launch");
}

bool vjp_supported(int d, int m, int disc_mode, int nblk) {
  if (m != 1 || d < 1 || d > 4) return false;
  if (disc_mode == PHYSS_DISC_GIVEN) return true;
  if (disc_mode != PHYSS_DISC_MATERN || nblk < 1 || d % nblk) return false;
  return true;                                                  // block size d / nblk in 1..4
}

int kf_vjp(cudaStream_t st, int d, int m, int disc_mode, int nblk, const SeqFilterArgs& a, const VjpOut& o) {
  if (!vjp_supported(d, m, disc_mode, nblk))
    return set_error(PHYSS_ERR_UNSUPPORTED, "filter vjp: supported for d <= 4, m == 1");
  if (disc_mode == PHYSS_DISC_GIVEN) {
    switch (d) {
      case 1: return vjp_launch<1, 1, true>(st, a, o);
      case 2: return vjp_launch<2, 2, true>(st, a, o);
      case 3: return vjp_launch<3, 3, true>(st, a, o);
      default: return vjp_launch<4, 4, true>(st, a, o);
    }
  }
  const int s = d / nblk;
  switch (d * 10 + s) {
    case 11: return vjp_launch<1, 1, false>(st, a, o);
    case 21: return vjp_launch<2, 1, false>(st, a, o);
    case 22: return vjp_launch<2, 2, false>(st, a, o);
    case 31: return vjp_launch<3, 1, false>(st, a, o);
    case 33: return vjp_launch<3, 3, false>(st, a, o);
    case 41: return vjp_launch<4, 1, false>(st, a, o);
    case 42: return vjp_launch<4, 2, false>(st, a, o);
    case 44: return vjp_launch<4, 4, false>(st, a, o);
    default: return set_error(PHYSS_ERR_UNSUPPORTED, "filter vjp: block size");
  }
}

}  // namespace physs
