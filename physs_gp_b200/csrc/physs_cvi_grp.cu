// physs_cvi_grp.cu -- CVI site update / expected log-likelihood for site blocks of size D > 4
// (derivative-augmented multi-latent states, BASELINE config 3): one lane group per (series, step)
// block, matrices in shared memory (physs_warp.cuh).  Same semantics as physs_cvi.cu.
#include "physs_cvi_core.cuh"
#include "physs_internal.h"
#include "physs_warp.cuh"

namespace physs {

using namespace grp;

struct CviGrpLayout {
  int D, P, ld, ldp;
  int M1, M2, qS, dS, WS, RW, Nz, Ri, W;
  int vY, vqm, vdm, vl1, vfmu, verr, vRe, vy, vrd, ve;
  int total;
};

static CviGrpLayout cvi_layout(int D, int P) {
  CviGrpLayout L{};
  L.D = D; L.P = P; L.ld = D | 1; L.ldp = P | 1;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  L.M1 = take(D * L.ld); L.M2 = take(D * L.ld); L.qS = take(D * L.ld); L.dS = take(D * L.ld);
  L.WS = take(P * L.ld); L.RW = take(P * L.ld); L.W = take(P * L.ld);
  L.Nz = take(P * L.ldp); L.Ri = take(P * L.ldp);
  L.vY = take(D); L.vqm = take(D); L.vdm = take(D); L.vl1 = take(D);
  L.vfmu = take(P); L.verr = take(P); L.vRe = take(P); L.vy = take(P); L.vrd = take(D > P ? D : P);
  L.ve = take(3 * P);
  L.total = off;
  return L;
}

// M <- (M + jit I)^-1 via Cholesky; `scratch` holds the factor.  Both [n x n] with leading dim ld.
template <int G>
__device__ __forceinline__ void spd_inverse_smem(double* M, double* scratch, int ld, int n, double jit,
                                                 double* rd) {
  const int gl = Lanes<G>::gl();
  for (int idx = gl; idx < n * n; idx += G) {
    const int i = idx / n, j = idx - i * n;
    scratch[i * ld + j] = M[i * ld + j] + (i == j ? jit : 0.0);
    M[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncwarp();
  chol<G>(scratch, ld, n, rd);
  chol_solve<G>(scratch, ld, n, rd, M, ld, n);
  __syncwarp();
}

template <int G, int LIK, bool UPDATE>
__global__ void cvi_grp_kernel(const CviArgs p, const CviGrpLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t n0 = (int64_t)blockIdx.x * gpb + g_in_block;
  const bool active = n0 < p.N;
  const int64_t n = active ? n0 : p.N - 1;
  const int gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int D = L.D, P = L.P, ld = L.ld, ldp = L.ldp;
  double* M1 = sm + L.M1; double* M2 = sm + L.M2; double* qS = sm + L.qS; double* dS = sm + L.dS;
  double* WS = sm + L.WS; double* RW = sm + L.RW; double* W = sm + L.W;
  double* Nz = sm + L.Nz; double* Ri = sm + L.Ri;
  double* Yt = sm + L.vY; double* qm = sm + L.vqm; double* dm = sm + L.vdm; double* l1 = sm + L.vl1;
  double* fmu = sm + L.vfmu; double* err = sm + L.verr; double* Re = sm + L.vRe; double* y = sm + L.vy;
  double* rd = sm + L.vrd; double* ev = sm + L.ve;

  g2s<G>(qS, ld, p.qS + n * D * D, D, D);
  for (int i = gl; i < D; i += G) qm[i] = p.qm[n * D + i];
  double ell = 0.0;
  if (LIK == CVI_LIK_GIVEN) {
    g2s<G>(dS, ld, p.dS_in + n * D * D, D, D);
    for (int i = gl; i < D; i += G) dm[i] = p.dm_in[n * D + i];
    __syncwarp();
  } else {
    for (int a = gl; a < P; a += G) y[a] = p.y[n * P + a];
    for (int idx = gl; idx < P * D; idx += G) {
      const int a = idx / D, k = idx - a * D;
      W[a * ld + k] = p.W ? p.W[idx] : (a == k ? 1.0 : 0.0);
    }
    __syncwarp();
    mv<G, false>(fmu, W, ld, qm, P, D, nullptr, 1.0);
    mm<G, false, false>(WS, ld, W, ld, qS, ld, P, D, D, nullptr, 0, 1.0);          // W S   [P x D]
    __syncwarp();
    if (LIK == CVI_LIK_GAUSS) {
      // masked noise -> Nz, identity -> Ri ; factor ; Ri <- R~^-1 ; restrict to observed
      const double* noise = p.noise + n * p.noise_stride;
      for (int idx = gl; idx < P * P; idx += G) {
        const int a = idx / P, b = idx - a * P;
        const bool keep = !(y[a] != y[a]) && !(y[b] != y[b]);
        Nz[a * ldp + b] = keep ? noise[idx] : (a == b ? 1.0 : 0.0);
        Ri[a * ldp + b] = (a == b) ? 1.0 : 0.0;
      }
      for (int a = gl; a < P; a += G) err[a] = (y[a] != y[a]) ? 0.0 : (y[a] - fmu[a]);
      __syncwarp();
      const double det = chol<G>(Nz, ldp, P, rd);
      chol_solve<G>(Nz, ldp, P, rd, Ri, ldp, P);
      __syncwarp();
      for (int idx = gl; idx < P * P; idx += G) {
        const int a = idx / P, b = idx - a * P;
        const bool keep = !(y[a] != y[a]) && !(y[b] != y[b]);
        if (!keep) Ri[a * ldp + b] = 0.0;
      }
      __syncwarp();
      mv<G, false>(Re, Ri, ldp, err, P, P, nullptr, 1.0);
      mm<G, false, false>(RW, ld, Ri, ldp, W, ld, P, P, D, nullptr, 0, 1.0);        // R^-1 W  [P x D]
      __syncwarp();
      // ell = -1/2 (nobs log 2pi + logdet + mahal + tr(R^-1 W S W^T))
      int nobs = 0;
      double mahal = 0.0, tr = 0.0;
      for (int a = 0; a < P; ++a) {
        nobs += (y[a] != y[a]) ? 0 : 1;
        mahal = fma(err[a], Re[a], mahal);
        for (int k = 0; k < D; ++k) tr = fma(RW[a * ld + k], WS[a * ld + k], tr);
      }
      ell = -0.5 * ((double)nobs * kLog2Pi + log(det) + mahal + tr);
      mv<G, true>(dm, W, ld, Re, D, P, nullptr, 1.0);                              // W^T R^-1 err
      mm<G, true, false>(dS, ld, W, ld, RW, ld, D, P, D, nullptr, 0, -0.5);        // -1/2 W^T R^-1 W
      __syncwarp();
    } else {
      // independent scalar sites, Gauss-Hermite; lane a handles output a
      for (int a = gl; a < P; a += G) {
        double fv = 0.0;
        for (int k = 0; k < D; ++k) fv = fma(WS[a * ld + k], W[a * ld + k], fv);
        const bool obs = !(y[a] != y[a]);
        const double ya = obs ? y[a] : 0.0;
        const double sd = sqrt(2.0 * fv);
        double e0 = 0.0, e1 = 0.0, e2 = 0.0;
        const double c0 = (LIK == CVI_LIK_POISSON_EXP) ? poisson_exp_const(ya, p.lik_param) : 0.0;
        for (int q = 0; q < p.K; ++q) {
          const double f = fma(sd, p.ghx[q], fmu[a]);
          double l, d1, d2;
          if (LIK == CVI_LIK_POISSON_EXP) poisson_exp_terms(ya, f, p.lik_param, c0, l, d1, d2);
          else bernoulli_probit_terms(ya, f, l, d1, d2);
          e0 = fma(p.ghw[q], l, e0);
          e1 = fma(p.ghw[q], d1, e1);
          e2 = fma(p.ghw[q], d2, e2);
        }
        ev[3 * a] = obs ? e0 : 0.0;
        ev[3 * a + 1] = obs ? e1 : 0.0;
        ev[3 * a + 2] = obs ? 0.5 * e2 : 0.0;
      }
      __syncwarp();
      for (int a = 0; a < P; ++a) ell += ev[3 * a];
      for (int i = gl; i < D; i += G) {
        double t = 0.0;
        for (int a = 0; a < P; ++a) t = fma(W[a * ld + i], ev[3 * a + 1], t);
        dm[i] = t;
      }
      for (int idx = gl; idx < D * D; idx += G) {
        const int i = idx / D, j = idx - i * D;
        double t = 0.0;
        for (int a = 0; a < P; ++a) t = fma(ev[3 * a + 2] * W[a * ld + i], W[a * ld + j], t);
        dS[i * ld + j] = t;
      }
      __syncwarp();
    }
    if (active) {
      if (p.ell && gl == 0) p.ell[n] = ell;
      if (p.dm_out) for (int i = gl; i < D; i += G) p.dm_out[n * D + i] = dm[i];
      if (p.dS_out) s2g<G>(p.dS_out + n * D * D, dS, ld, D, D);
    }
  }
  if (UPDATE) {
    g2s<G>(M1, ld, p.Vt + n * D * D, D, D);
    for (int i = gl; i < D; i += G) Yt[i] = p.Yt[n * D + i];
    __syncwarp();
    spd_inverse_smem<G>(M1, M2, ld, D, p.ngj, rd);                   // M1 = (V~ + ngj I)^-1
    // lambda_1' and P = -2 lambda_2'  (into M2)
    for (int i = gl; i < D; i += G) {
      double t = 0.0, g = dm[i];
      for (int k = 0; k < D; ++k) {
        t = fma(M1[i * ld + k], Yt[k], t);
        g = fma(-2.0 * dS[i * ld + k], qm[k], g);
      }
      l1[i] = (1.0 - p.beta) * t + p.beta * g;
    }
    __syncwarp();
    // 'NG_Precision' (p.prec): the site matrix IS the precision -- lambda_2 = -1/2 of it (re-read: M1 now holds the
    // inverse that lambda_1 uses, as the reference's cholesky_solve does) and theta_2' = -2 lambda_2' goes out as it is
    for (int idx = gl; idx < D * D; idx += G) {
      const int i = idx / D, j = idx - i * D;
      const double src = p.prec ? p.Vt[n * D * D + idx] : M1[i * ld + j];
      const double l2 = (1.0 - p.beta) * (-0.5 * src) + p.beta * dS[i * ld + j];
      M2[i * ld + j] = -2.0 * l2;
    }
    __syncwarp();
    if (p.prec && active) s2g<G>(p.Vn + n * D * D, M2, ld, D, D);
    __syncwarp();
    spd_inverse_smem<G>(M2, M1, ld, D, p.ngj, rd);                   // M2 = V~'  (prec: the inverse theta_1' needs)
    if (active) {
      for (int i = gl; i < D; i += G) {
        double t = 0.0;
        for (int k = 0; k < D; ++k) t = fma(M2[i * ld + k], l1[k], t);
        p.Yn[n * D + i] = t;
      }
      if (!p.prec) s2g<G>(p.Vn + n * D * D, M2, ld, D, D);
    }
  }
}

template <int G, int LIK, bool UPDATE>
static int run(cudaStream_t st, const CviArgs& a, const CviGrpLayout& L) {
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 200 * 1024) threads /= 2;
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024) return set_error(PHYSS_ERR_UNSUPPORTED, "cvi: site block too large");
  const int gpb = threads / G;
  const int64_t grid = (a.N + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(cvi_grp_kernel<G, LIK, UPDATE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(cvi_grp_kernel)");
  cvi_grp_kernel<G, LIK, UPDATE><<<(unsigned)grid, threads, smem, st>>>(a, L);
  return cuda_status(cudaGetLastError(), "cvi_grp_kernel launch");
}

template <int G, bool UPDATE>
static int by_lik(cudaStream_t st, const CviArgs& a, const CviGrpLayout& L, int lik) {
  switch (lik) {
    case CVI_LIK_GAUSS: return run<G, CVI_LIK_GAUSS, UPDATE>(st, a, L);
    case CVI_LIK_POISSON_EXP: return run<G, CVI_LIK_POISSON_EXP, UPDATE>(st, a, L);
    case CVI_LIK_BERNOULLI_PROBIT: return run<G, CVI_LIK_BERNOULLI_PROBIT, UPDATE>(st, a, L);
    case CVI_LIK_GIVEN:
      if (UPDATE) return run<G, CVI_LIK_GIVEN, UPDATE>(st, a, L);
      break;
  }
  return set_error(PHYSS_ERR_BAD_ARG, "cvi: unknown likelihood kind");
}

int cvi_grp_run(cudaStream_t st, int D, int P, int lik, bool update, const CviArgs& a) {
  if (D < 1 || D > 64 || P < 1 || P > D) return set_error(PHYSS_ERR_UNSUPPORTED, "cvi: need 1 <= P <= D <= 64");
  const CviGrpLayout L = cvi_layout(D, lik == CVI_LIK_GIVEN ? 1 : P);
  const int G = D <= 8 ? 8 : (D <= 16 ? 16 : 32);
#define RUN(G_) (update ? by_lik<G_, true>(st, a, L, lik) : by_lik<G_, false>(st, a, L, lik))
  if (G == 8) return RUN(8);
  if (G == 16) return RUN(16);
  return RUN(32);
#undef RUN
}

}  // namespace physs
