// physs_pscan.cu -- parallel-in-time Kalman filter / RTS smoother: chunked associative scan.
//
// Replaces filter('parallel') / smoother('parallel') (computation/filters/parallel_kalman_filter.py:
// 225-336, parallel_rts_smoother.py:57-103).  The reference materialises one scan element per time step
// and runs jax.lax.associative_scan over all T of them (~2T combines of ~19 d^3 flop, log2 T full passes
// over HBM).  Here the time axis is cut into chunks of `chunk_len` steps and the same associative
// operators are used BETWEEN chunks only:
//
//   1. summary   one lane group per (series, chunk) folds its steps into ONE element, sequentially and in
//                shared memory.  Folding a single-step element into an accumulated one collapses
//                algebraically to a Kalman predict/update on (b, C) plus rank-m updates of (A, J, eta)
//                (filter: element (A, b, C, J, eta) of parallel_kalman_filter.py:143-162, operator
//                :178-220), or an RTS step on (g, L) plus E <- G E (smoother: element (E, g, L) of
//                parallel_rts_smoother.py:25-37, operator :39-55).  No per-step element touches HBM.
//   2. scan      Hillis-Steele inclusive scan over the chunk summaries with the generic operators
//                (general LU solves with partial pivoting, as jsp.linalg.solve(assume_a='gen')),
//                log2(nchunk) launches over B * nchunk elements.
//   3. apply     the start state (prior, or the carried state of the previous time shard on another GPU)
//                is pushed through every prefix -> the exact state at every chunk boundary.
//   4. replay    the ordinary sequential kernels (physs_seq*.cu / physs_grp.cu, chunk mode) run all chunks
//                concurrently from their boundary states and write the per-step moments.
//   5. polish    (filter, jitter != 0 only) the reference's sequential recursion adds `jitter` inside the
//                gain solve but not in P - K S K^T, which no scan element can represent
//                exactly; boundary states from the scan are therefore O(jitter) away from the sequential
//                recursion the parity oracle runs.  Each chunk is restarted from the previous chunk's
//                replayed end state and re-run until it agrees with what is stored (contraction of the
//                filter); `unconverged` is raised if a chunk reaches its end without agreeing.
//
// Elements in global memory (per (series b, chunk c), index v = b * nchunk + c):
//   filter   [A d*d | C d*d | J d*d | b d | eta d]      smoother   [E d*d | L d*d | g d]
#include <stdlib.h>

#include <cuda/ptx>

#include "physs_internal.h"
#include "physs_warp.cuh"

namespace physs {

using namespace grp;

static inline int ps_group_size(int d) { return d <= 8 ? 8 : (d <= 16 ? 16 : 32); }
// PHYSS_FORCE_GRP=1 (A-B timing only): runtime-sized summary kernels instead of the register-tiled ones
static bool ps_force_grp() {
  static const bool on = [] { const char* e = getenv("PHYSS_FORCE_GRP"); return e && e[0] == '1'; }();
  return on;
}

// ------------------------------------------------------------------- TMA (bulk-copy) staging of scan elements
// A scan element is a handful of dense row-major blocks in global memory ([A | C | J | b | eta] or
// [E | L | g]).  For even d every row is a multiple of 16 bytes and 16-byte aligned, so ONE elected lane of
// the group stages the whole element with 1-D bulk copies (cp.async.bulk.shared.global, the TMA engine) into
// the padded shared-memory rows and arms the group's mbarrier with the byte count; the other lanes only wait
// on the barrier.  Odd d falls back to per-lane loads.
namespace ptx = cuda::ptx;

__device__ __forceinline__ void tma_rows(double* sdst, int ld, const double* __restrict__ gsrc, int n, int m,
                                         uint64_t* bar) {
  for (int i = 0; i < n; ++i)
    ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, sdst + i * ld, gsrc + (size_t)i * m,
                       (uint32_t)(m * sizeof(double)), bar);
}
// bounded spin on the mbarrier phase (a lost completion must surface as a wrong result, not as a hung GPU)
__device__ __forceinline__ void tma_wait(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 20); ++it)
    if (ptx::mbarrier_try_wait_parity(bar, parity)) return;
}
template <int G>
__device__ __forceinline__ void tma_bar_init(uint64_t* bar) {
  if (Lanes<G>::gl() == 0) {
    ptx::mbarrier_init(bar, 1);
    ptx::fence_proxy_async(ptx::space_shared);      // make the initialised barrier visible to the async proxy
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------ layouts
struct PsLayout {
  int d, m, ld, ldm, nblk, s;
  int C, A, Qm, W1, W2, Acc, Abar, GE, J, HA, Z, S, Sj, H, Rst[2], AQst[2][2], PfS[2];
  int vb, vbb, veta, vv, vw, vrd, vy[2], vmf[2], vlam, vdm;
  int total;
};

static PsLayout ps_layout(int d, int m, int nblk, bool given, bool smoother) {
  PsLayout L{};
  L.d = d; L.m = m; L.ld = d | 1; L.ldm = (m > 0 ? m : 1) | 1; L.nblk = nblk;
  L.s = (nblk > 0) ? d / nblk : d;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  const int dd = d * L.ld;
  L.C = take(dd); L.A = take(dd); L.Qm = take(dd); L.W1 = take(dd); L.W2 = take(dd);
  L.Acc = take(dd); L.Abar = take(dd);
  if (smoother) {
    L.PfS[0] = take(dd); L.PfS[1] = take(dd); L.GE = take(dd);
    L.vmf[0] = take(d); L.vmf[1] = take(d);
  } else {
    L.J = take(dd);
    L.HA = take(m * L.ld); L.Z = take(m * L.ld);
    L.S = take(m * L.ldm); L.Sj = take(m * L.ldm);
    L.H = take(m * L.ld);
    L.Rst[0] = take(m * L.ldm); L.Rst[1] = take(m * L.ldm);
    L.vy[0] = take(m); L.vy[1] = take(m);
    L.vv = take(m); L.vw = take(m);
    L.veta = take(d);
  }
  if (given) {
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L.AQst[a][b] = take(dd);
  }
  L.vb = take(d); L.vbb = take(d); L.vdm = take(d); L.vrd = take(d > m ? d : m);
  L.vlam = take(nblk > 0 ? nblk : 1);
  L.total = off;
  return L;
}

__host__ __device__ inline int64_t ps_filter_elem(int d) { return 3LL * d * d + 2 * d; }
__host__ __device__ inline int64_t ps_smooth_elem(int d) { return 2LL * d * d + d; }

// ------------------------------------------------------------------------------- 1. filter summary
// One group per (b, c), c in [cfirst, cfirst + ccount).  All chunks of one launch have the same length
// (uniform control flow inside a warp); the ragged tail chunk gets its own launch.
template <int G, bool GIVEN>
__global__ void ps_filter_summary_kernel(const SeqFilterArgs p, const PsLayout L, const bool hid,
                                         const int64_t cfirst, const int64_t ccount,
                                         double* __restrict__ elems) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = p.B * ccount;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t bb = g / ccount, c = cfirst + g % ccount;
  const int64_t t0 = c * p.chunk_len;
  const int64_t T = (p.chunk_len < p.T - t0) ? p.chunk_len : (p.T - t0);
  const int gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = L.d, m = L.m, ld = L.ld, ldm = L.ldm, s = L.s;

  double* C = sm + L.C; double* A = sm + L.A; double* Qm = sm + L.Qm;
  double* W1 = sm + L.W1; double* W2 = sm + L.W2;
  double* Acc = sm + L.Acc; double* Abar = sm + L.Abar; double* J = sm + L.J;
  double* HA = sm + L.HA; double* Z = sm + L.Z;
  double* S = sm + L.S; double* Sj = sm + L.Sj; double* H = sm + L.H;
  double* bv = sm + L.vb; double* bbar = sm + L.vbb; double* eta = sm + L.veta;
  double* v = sm + L.vv; double* w = sm + L.vw; double* rd = sm + L.vrd; double* lam = sm + L.vlam;

  // conditional element of an empty interval: A = I, b = 0, C = 0, J = 0, eta = 0
  for (int idx = gl; idx < d * d; idx += G) {
    const int i = idx / d, j = idx - i * d;
    Acc[i * ld + j] = (i == j) ? 1.0 : 0.0;
    C[i * ld + j] = 0.0;
    J[i * ld + j] = 0.0;
  }
  for (int i = gl; i < d; i += G) { bv[i] = 0.0; eta[i] = 0.0; }
  if (!GIVEN) {
    g2s<G>(Qm, ld, p.Pinf + bb * p.Pinf_bs, d, d);
    for (int i = gl; i < L.nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
  }
  if (!hid) g2s<G>(H, ld, p.H + bb * p.H_bs, m, d);
  const double* dtp = p.dt + bb * p.dt_bs + t0;
  const int64_t sts = p.sts;
  const int64_t row0 = bb * p.sbs + t0 * sts;
  const double* Yp = p.Y + row0 * m;
  const double* Rp = p.R + bb * p.R_bs + t0 * p.R_ts;
  const double* Ap = GIVEN ? p.A + bb * p.A_bs + t0 * d * d : nullptr;
  const double* Qp = GIVEN ? p.Q + bb * p.Q_bs + t0 * d * d : nullptr;

  auto stage = [&](int64_t k) {
    const int st = (int)(k & 1);
    g2s_async<G>(sm + L.vy[st], m, Yp + k * sts * m, 1, m);
    g2s_async<G>(sm + L.Rst[st], ldm, Rp + k * p.R_ts, m, m);
    if (GIVEN) {
      g2s_async<G>(sm + L.AQst[st][0], ld, Ap + k * d * d, d, d);
      g2s_async<G>(sm + L.AQst[st][1], ld, Qp + k * d * d, d, d);
    }
    cp_async_commit();
  };

  stage(0);
  double dt_n = dtp[0];
  __syncwarp();
  for (int64_t k = 0; k < T; ++k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    cp_async_wait_all();
    __syncwarp();
    if (k + 1 < T) { stage(k + 1); dt_n = dtp[k + 1]; }
    const double* y = sm + L.vy[st];
    const double* R = sm + L.Rst[st];
    // ---- predict (b, C) and push the transition through A:  Abar = Phi Acc
    if (GIVEN) {
      const double* Ak = sm + L.AQst[st][0];
      const double* Qk = sm + L.AQst[st][1];
      mv<G, false>(bbar, Ak, ld, bv, d, d, nullptr, 1.0);
      mm<G, false, false>(W2, ld, Ak, ld, C, ld, d, d, d, nullptr, 0, 1.0);
      mm<G, false, false>(Abar, ld, Ak, ld, Acc, ld, d, d, d, nullptr, 0, 1.0);
      __syncwarp();
      mm<G, false, true>(C, ld, W2, ld, Ak, ld, d, d, d, Qk, ld, 1.0);
    } else {
      matern_A<G>(A, ld, d, s, L.nblk, lam, dt);
      for (int idx = gl; idx < d * d; idx += G) {
        const int i = idx / d, j = idx - i * d;
        W1[i * ld + j] = C[i * ld + j] - Qm[i * ld + j];
      }
      __syncwarp();
      mv<G, false>(bbar, A, ld, bv, d, d, nullptr, 1.0, s);
      mm<G, false, false>(W2, ld, A, ld, W1, ld, d, d, d, nullptr, 0, 1.0, s, 0);
      mm<G, false, false>(Abar, ld, A, ld, Acc, ld, d, d, d, nullptr, 0, 1.0, s, 0);
      __syncwarp();
      mm<G, false, true>(C, ld, W2, ld, A, ld, d, d, d, Qm, ld, 1.0, 0, s);
    }
    __syncwarp();
    // ---- masked update (kalman_filter.py:144-211): W1 = M H C_ [m x d], HA = M H Abar [m x d]
    for (int idx = gl; idx < m * d; idx += G) {
      const int a = idx / d, j = idx - a * d;
      double hc, ha;
      if (hid) {
        hc = C[a * ld + j];
        ha = Abar[a * ld + j];
      } else {
        hc = 0.0; ha = 0.0;
        for (int l = 0; l < d; ++l) {
          hc = fma(H[a * ld + l], C[l * ld + j], hc);
          ha = fma(H[a * ld + l], Abar[l * ld + j], ha);
        }
      }
      const double ya = y[a];
      const bool miss = (ya != ya);
      W1[a * ld + j] = miss ? 0.0 : hc;
      HA[a * ld + j] = miss ? 0.0 : ha;
      Z[a * ld + j] = miss ? 0.0 : ha;
    }
    for (int a = gl; a < m; a += G) {
      double mu;
      if (hid) {
        mu = bbar[a];
      } else {
        mu = 0.0;
        for (int l = 0; l < d; ++l) mu = fma(H[a * ld + l], bbar[l], mu);
      }
      const double ya = y[a];
      const double r = (ya != ya) ? 0.0 : (ya - mu);
      v[a] = r;
      w[a] = r;
    }
    __syncwarp();
    for (int idx = gl; idx < m * m; idx += G) {
      const int a = idx / m, cc = idx - a * m;
      double accv;
      if (hid) {
        accv = W1[a * ld + cc];
      } else {
        accv = 0.0;
        for (int l = 0; l < d; ++l) accv = fma(W1[a * ld + l], H[cc * ld + l], accv);
      }
      const double yc = y[cc];
      accv = (yc != yc) ? 0.0 : accv;
      const double sv = accv + R[a * ldm + cc];
      S[a * ldm + cc] = sv;
      Sj[a * ldm + cc] = sv + (a == cc ? p.jitter : 0.0);
    }
    __syncwarp();
    chol<G>(Sj, ldm, m, rd);
    chol_solve<G>(Sj, ldm, m, rd, W1, ld, d);       // W1 <- K^T
    chol_solve<G>(Sj, ldm, m, rd, Z, ld, d);        // Z  <- (S + jit)^-1 M H Abar
    chol_solve<G>(Sj, ldm, m, rd, w, 1, 1);         // w  <- (S + jit)^-1 v
    __syncwarp();
    // b = bbar + K v ; eta += HA^T w
    for (int i = gl; i < d; i += G) {
      double accb = bbar[i], acce = eta[i];
      for (int a = 0; a < m; ++a) {
        accb = fma(W1[a * ld + i], v[a], accb);
        acce = fma(HA[a * ld + i], w[a], acce);
      }
      bv[i] = accb;
      eta[i] = acce;
    }
    // KS = K S -> W2 [d x m]
    for (int idx = gl; idx < d * m; idx += G) {
      const int i = idx / m, cc = idx - i * m;
      double accv = 0.0;
      for (int a = 0; a < m; ++a) accv = fma(W1[a * ld + i], S[a * ldm + cc], accv);
      W2[i * ldm + cc] = accv;
    }
    __syncwarp();
    // C -= KS K^T ; Acc = Abar - K HA ; J += HA^T Z
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      double cv = C[i * ld + j], av = Abar[i * ld + j], jv = J[i * ld + j];
      for (int a = 0; a < m; ++a) {
        cv = fma(-W2[i * ldm + a], W1[a * ld + j], cv);
        av = fma(-W1[a * ld + i], HA[a * ld + j], av);
        jv = fma(HA[a * ld + i], Z[a * ld + j], jv);
      }
      C[i * ld + j] = cv;
      Acc[i * ld + j] = av;
      J[i * ld + j] = jv;
    }
    __syncwarp();
  }
  if (active) {
    double* e = elems + (bb * p.nchunk + c) * ps_filter_elem(d);
    s2g<G>(e, Acc, ld, d, d);
    s2g<G>(e + d * d, C, ld, d, d);
    s2g<G>(e + 2 * d * d, J, ld, d, d);
    for (int i = gl; i < d; i += G) { e[3 * d * d + i] = bv[i]; e[3 * d * d + d + i] = eta[i]; }
  }
}

// ------------------------------------------------------------------------------ 2. filter combine
struct PcLayout {
  int d, ld;
  int Ai, Ci, Ji, Aj, Cj, Jj, M1, M1T, X1, X2, W, bi, ei, bj, ej, t1, t2, bar;
  int total;
};
static PcLayout pc_layout(int d) {
  PcLayout L{};
  L.d = d; L.ld = (d % 2 == 0) ? d + 2 : (d | 1);      // even d: 16-byte aligned rows (TMA staging)
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  const int dd = d * L.ld;
  L.Ai = take(dd); L.Ci = take(dd); L.Ji = take(dd); L.Aj = take(dd); L.Cj = take(dd); L.Jj = take(dd);
  L.M1 = take(dd); L.M1T = take(dd); L.X1 = take(dd); L.X2 = take(dd); L.W = take(dd);
  L.bi = take(d); L.ei = take(d); L.bj = take(d); L.ej = take(d); L.t1 = take(d); L.t2 = take(d);
  L.bar = take(2);
  L.total = off;
  return L;
}

// (A, b, C, J, eta)_out = left (x) right, filtering_operator of parallel_kalman_filter.py:178-220 with the
// general solves of its default branch (:201-211) and force_symmetric on C, J (:216-219).
// Operands in shared memory; the result overwrites the `j` slots (Aj, Cj, Jj, bj, ej).
// DC != 0: the state dimension as a compile-time constant (d = 6, 8, 12: the physics-informed shapes of config 3) --
// the index divisions, loop bounds and row strides of the generic products and LU solves below then fold to constants
// (the runtime-sized form spent 100 us on ONE 8 x 8 combine, 1.2 of the 10 ms of a 1M-step scan).  Same arithmetic in
// the same order: bitwise equal to DC = 0.
template <int G, int DC = 0>
__device__ __forceinline__ void filter_combine(double* sm, const PcLayout& L) {
  const int gl = Lanes<G>::gl();
  const int d = DC ? DC : L.d, ld = DC ? ((DC % 2 == 0) ? DC + 2 : (DC | 1)) : L.ld;
  double* Ai = sm + L.Ai; double* Ci = sm + L.Ci; double* Ji = sm + L.Ji;
  double* Aj = sm + L.Aj; double* Cj = sm + L.Cj; double* Jj = sm + L.Jj;
  double* M1 = sm + L.M1; double* M1T = sm + L.M1T; double* X1 = sm + L.X1; double* X2 = sm + L.X2;
  double* W = sm + L.W;
  double* bi = sm + L.bi; double* ei = sm + L.ei; double* bj = sm + L.bj; double* ej = sm + L.ej;
  double* t1 = sm + L.t1; double* t2 = sm + L.t2;
  // M1 = I + Ci Jj ; M1T = M1^T ; X1 = Aj^T ; X2 = Ai
  for (int idx = gl; idx < d * d; idx += G) {
    const int i = idx / d, j = idx - i * d;
    double acc = (i == j) ? 1.0 : 0.0;
    for (int l = 0; l < d; ++l) acc = fma(Ci[i * ld + l], Jj[l * ld + j], acc);
    M1[i * ld + j] = acc;
    M1T[j * ld + i] = acc;
    X1[i * ld + j] = Aj[j * ld + i];
    X2[i * ld + j] = Ai[i * ld + j];
  }
  // t1 = bi + Ci ej ; t2 = ej - Jj bi
  for (int i = gl; i < d; i += G) {
    double a1 = bi[i], a2 = ej[i];
    for (int l = 0; l < d; ++l) {
      a1 = fma(Ci[i * ld + l], ej[l], a1);
      a2 = fma(-Jj[i * ld + l], bi[l], a2);
    }
    t1[i] = a1;
    t2[i] = a2;
  }
  __syncwarp();
  lu_solve<G>(M1T, ld, d, X1, ld, d);      // X1 = M1^-T Aj^T = (Aj M1^-1)^T
  __syncwarp();
  lu_solve<G>(M1, ld, d, X2, ld, d);       // X2 = M1^-1 Ai   ( = (Ai^T (I + Jj Ci)^-1)^T )
  __syncwarp();
  // W = X1^T Ci (= Aj_tmp Ci)
  mm<G, true, false>(W, ld, X1, ld, Ci, ld, d, d, d, nullptr, 0, 1.0);
  // M1 <- A_out = X1^T Ai
  mm<G, true, false>(M1, ld, X1, ld, Ai, ld, d, d, d, nullptr, 0, 1.0);
  // M1T <- X2^T Jj (= Ai_tmp Jj)
  mm<G, true, false>(M1T, ld, X2, ld, Jj, ld, d, d, d, nullptr, 0, 1.0);
  // b_out = X1^T t1 + bj ; eta_out = X2^T t2 + ei      (write to bj / ej after the reads above)
  __syncwarp();
  for (int i = gl; i < d; i += G) {
    double a1 = bj[i], a2 = ei[i];
    for (int l = 0; l < d; ++l) {
      a1 = fma(X1[l * ld + i], t1[l], a1);
      a2 = fma(X2[l * ld + i], t2[l], a2);
    }
    bj[i] = a1;
    ej[i] = a2;
  }
  // X1 <- C_out = W Aj^T + Cj ; X2 <- J_out = (X2^T Jj) Ai + Ji
  mm<G, false, true>(X1, ld, W, ld, Aj, ld, d, d, d, Cj, ld, 1.0);
  mm<G, false, false>(X2, ld, M1T, ld, Ai, ld, d, d, d, Ji, ld, 1.0);
  __syncwarp();
  for (int idx = gl; idx < d * d; idx += G) {
    const int i = idx / d, j = idx - i * d;
    Aj[i * ld + j] = M1[i * ld + j];
    Cj[i * ld + j] = 0.5 * (X1[i * ld + j] + X1[j * ld + i]);
    Jj[i * ld + j] = 0.5 * (X2[i * ld + j] + X2[j * ld + i]);
  }
  __syncwarp();
}

// issue (no wait): element e -> slot `right ? j : i`.  Returns the bytes in flight on the barrier (0 = the
// copy was done synchronously with per-lane loads).
template <int G>
__device__ __forceinline__ uint32_t load_filter_elem(double* sm, const PcLayout& L, bool right,
                                                     const double* __restrict__ e) {
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  double* A = sm + (right ? L.Aj : L.Ai); double* C = sm + (right ? L.Cj : L.Ci); double* J = sm + (right ? L.Jj : L.Ji);
  double* b = sm + (right ? L.bj : L.bi); double* eta = sm + (right ? L.ej : L.ei);
  if ((d & 1) == 0) {
    if (gl == 0) {
      uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
      ptx::fence_proxy_async(ptx::space_shared);   // order earlier generic-proxy accesses to these rows
      tma_rows(A, ld, e, d, d, bar);
      tma_rows(C, ld, e + d * d, d, d, bar);
      tma_rows(J, ld, e + 2 * d * d, d, d, bar);
      tma_rows(b, d, e + 3 * d * d, 1, d, bar);
      tma_rows(eta, d, e + 3 * d * d + d, 1, d, bar);
    }
    return (uint32_t)((3 * d * d + 2 * d) * sizeof(double));
  }
  g2s<G>(A, ld, e, d, d);
  g2s<G>(C, ld, e + d * d, d, d);
  g2s<G>(J, ld, e + 2 * d * d, d, d);
  for (int i = gl; i < d; i += G) { b[i] = e[3 * d * d + i]; eta[i] = e[3 * d * d + d + i]; }
  return 0;
}
// arm the barrier with the bytes issued by load_*_elem and wait for this phase
template <int G>
__device__ __forceinline__ void tma_finish(double* sm, int bar_off, uint32_t bytes, uint32_t& parity) {
  if (bytes == 0) { __syncwarp(); return; }
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + bar_off);
  if (Lanes<G>::gl() == 0) ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, bar, bytes);
  __syncwarp();                                     // the arrival is posted before anybody starts to poll
  tma_wait(bar, parity);
  parity ^= 1u;
  __syncwarp();
}
// left operand = identity element (A = I, rest 0) or a bare state (A = 0, b = m, C = P, J = 0, eta = 0)
template <int G>
__device__ __forceinline__ void left_special(double* sm, const PcLayout& L, const double* __restrict__ mstate,
                                             const double* __restrict__ Pstate) {
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  for (int idx = gl; idx < d * d; idx += G) {
    const int i = idx / d, j = idx - i * d;
    sm[L.Ai + i * ld + j] = (!Pstate && i == j) ? 1.0 : 0.0;
    sm[L.Ci + i * ld + j] = Pstate ? Pstate[idx] : 0.0;
    sm[L.Ji + i * ld + j] = 0.0;
  }
  for (int i = gl; i < d; i += G) {
    sm[L.bi + i] = mstate ? mstate[i] : 0.0;
    sm[L.ei + i] = 0.0;
  }
}
template <int G>
__device__ __forceinline__ void store_filter_elem(const double* sm, const PcLayout& L, double* __restrict__ e) {
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  s2g<G>(e, sm + L.Aj, ld, d, d);
  s2g<G>(e + d * d, sm + L.Cj, ld, d, d);
  s2g<G>(e + 2 * d * d, sm + L.Jj, ld, d, d);
  for (int i = gl; i < d; i += G) { e[3 * d * d + i] = sm[L.bj + i]; e[3 * d * d + d + i] = sm[L.ej + i]; }
}

// Hillis-Steele step: out[c] = in[c - stride] (x) in[c]  (identity on the left for c < stride)
template <int G, int DC = 0>
__global__ void ps_filter_scan_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t B,
                                      int64_t nchunk, int64_t nsum, int64_t stride, const PcLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = B * nsum;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t b = g / nsum, c = g % nsum;
  double* sm = smem + (size_t)g_in_block * L.total;
  const int64_t ne = ps_filter_elem(L.d);
  tma_bar_init<G>(reinterpret_cast<uint64_t*>(sm + L.bar));
  uint32_t parity = 0;
  uint32_t bytes = load_filter_elem<G>(sm, L, true, in + (b * nchunk + c) * ne);
  if (c >= stride) bytes += load_filter_elem<G>(sm, L, false, in + (b * nchunk + c - stride) * ne);
  else left_special<G>(sm, L, nullptr, nullptr);
  tma_finish<G>(sm, L.bar, bytes, parity);
  filter_combine<G, DC>(sm, L);
  if (active) store_filter_elem<G>(sm, L, out + (b * nchunk + c) * ne);
}

// 3. apply: boundary state of chunk c + 1 = (b, C) of  [start state] (x) prefix[c];  boundary 0 = start.
// start state: (m0, P0) per series (strides), or start_m / start_P [B, d] / [B, d, d] when given.
// Also used with nsum == 1 and `tot_out`: fold ONE element (the total of a previous time shard) onto a
// state, writing the new state (multi-GPU carry).
template <int G>
__global__ void ps_filter_apply_kernel(const double* __restrict__ prefix, int64_t B, int64_t nchunk,
                                       int64_t nsum, const double* __restrict__ m0, int64_t m0_bs,
                                       const double* __restrict__ P0, int64_t P0_bs,
                                       double* __restrict__ bnd_m, double* __restrict__ bnd_P,
                                       const PcLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = B * nsum;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t b = g / nsum, c = g % nsum;
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int64_t ne = ps_filter_elem(d);
  tma_bar_init<G>(reinterpret_cast<uint64_t*>(sm + L.bar));
  uint32_t parity = 0;
  const uint32_t bytes = load_filter_elem<G>(sm, L, true, prefix + (b * nchunk + c) * ne);
  left_special<G>(sm, L, m0 + b * m0_bs, P0 + b * P0_bs);
  tma_finish<G>(sm, L.bar, bytes, parity);
  filter_combine<G>(sm, L);
  if (active) {
    // bnd index c + 1 (< nchunk by construction of the callers); boundary 0 is the start state itself
    double* om = bnd_m + (b * nchunk + c + 1) * d;
    double* oP = bnd_P + (b * nchunk + c + 1) * d * d;
    for (int i = gl; i < d; i += G) om[i] = sm[L.bj + i];
    s2g<G>(oP, sm + L.Cj, ld, d, d);
    if (c == 0) {
      for (int i = gl; i < d; i += G) bnd_m[(b * nchunk) * d + i] = m0[b * m0_bs + i];
      for (int idx = gl; idx < d * d; idx += G) bnd_P[(b * nchunk) * d * d + idx] = P0[b * P0_bs + idx];
    }
  }
}

// ----------------------------------------------------------------------------- smoother summary
// One group per (b, c), c in [0, nchunk): folds the RTS steps of the chunk, last step first, into
// (E, g, L): x_s[t0] = E x + g, P_s[t0] = E P E^T + L for the smoothed state (x, P) of step t0 + len.
template <int G, bool GIVEN>
__global__ void ps_smooth_summary_kernel(const SeqSmoothArgs p, const PsLayout L, double* __restrict__ elems) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = p.B * p.chunk_count;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t bb = g / p.chunk_count, c = p.chunk_first + g % p.chunk_count;
  const int64_t t0 = c * p.chunk_len;
  const int64_t T = (p.chunk_len < p.T - t0) ? p.chunk_len : (p.T - t0);
  const int gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = L.d, ld = L.ld, s = L.s;

  double* Ls = sm + L.C; double* A = sm + L.A; double* Qm = sm + L.Qm;
  double* W1 = sm + L.W1; double* W2 = sm + L.W2; double* E = sm + L.Acc; double* W3 = sm + L.Abar;
  double* GE = sm + L.GE;
  double* gv = sm + L.vb; double* mpred = sm + L.vbb; double* dm = sm + L.vdm;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;

  for (int idx = gl; idx < d * d; idx += G) {
    const int i = idx / d, j = idx - i * d;
    E[i * ld + j] = (i == j) ? 1.0 : 0.0;
    Ls[i * ld + j] = 0.0;
  }
  for (int i = gl; i < d; i += G) gv[i] = 0.0;
  if (!GIVEN) {
    g2s<G>(Qm, ld, p.Pinf + bb * p.Pinf_bs, d, d);
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
    g2s_async<G>(sm + L.vmf[st], d, mfp + k * sts * d, 1, d);
    g2s_async<G>(sm + L.PfS[st], ld, Pfp + k * sts * d * d, d, d);
    if (GIVEN) {
      g2s_async<G>(sm + L.AQst[st][0], ld, Ap + k * d * d, d, d);
      g2s_async<G>(sm + L.AQst[st][1], ld, Qp + k * d * d, d, d);
    }
    cp_async_commit();
  };

  stage(T - 1);
  double dt_n = dtp[T - 1];
  __syncwarp();
  for (int64_t k = T - 1; k >= 0; --k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    cp_async_wait_all();
    __syncwarp();
    if (k >= 1) { stage(k - 1); dt_n = dtp[k - 1]; }
    const double* mf = sm + L.vmf[st];
    const double* Pf = sm + L.PfS[st];
    if (GIVEN) {
      const double* Ak = sm + L.AQst[st][0];
      const double* Qk = sm + L.AQst[st][1];
      mv<G, false>(mpred, Ak, ld, mf, d, d, nullptr, 1.0);
      mm<G, false, false>(W1, ld, Ak, ld, Pf, ld, d, d, d, nullptr, 0, 1.0);
      __syncwarp();
      mm<G, false, true>(W2, ld, W1, ld, Ak, ld, d, d, d, Qk, ld, 1.0);
    } else {
      matern_A<G>(A, ld, d, s, L.nblk, lam, dt);
      __syncwarp();
      mv<G, false>(mpred, A, ld, mf, d, d, nullptr, 1.0, s);
      mm<G, false, false>(W1, ld, A, ld, Pf, ld, d, d, d, nullptr, 0, 1.0, s, 0);
      mm<G, false, false>(W3, ld, A, ld, Qm, ld, d, d, d, nullptr, 0, 1.0, s, 0);
      __syncwarp();
      for (int idx = gl; idx < d * d; idx += G) {
        const int i = idx / d, j = idx - i * d;
        W3[i * ld + j] = W1[i * ld + j] - W3[i * ld + j];
      }
      __syncwarp();
      mm<G, false, true>(W2, ld, W3, ld, A, ld, d, d, d, Qm, ld, 1.0, 0, s);
    }
    __syncwarp();
    // dL = L - Pp -> W3 ; dm = g - mpred ; factor Pp + jitter I
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      const double pp = W2[i * ld + j];
      W3[i * ld + j] = Ls[i * ld + j] - pp;
      if (i == j) W2[i * ld + j] = pp + p.jitter;
    }
    for (int i = gl; i < d; i += G) dm[i] = gv[i] - mpred[i];
    __syncwarp();
    chol<G>(W2, ld, d, rd);
    chol_solve<G>(W2, ld, d, rd, W1, ld, d);                 // W1 <- X = G^T
    __syncwarp();
    mv<G, true>(gv, W1, ld, dm, d, d, mf, 1.0);                // g = mf + G (g - mpred)
    mm<G, true, false>(W2, ld, W1, ld, W3, ld, d, d, d, nullptr, 0, 1.0);   // G dL
    mm<G, true, false>(GE, ld, W1, ld, E, ld, d, d, d, nullptr, 0, 1.0);    // G E
    __syncwarp();
    mm<G, false, false>(Ls, ld, W2, ld, W1, ld, d, d, d, Pf, ld, 1.0);      // L = Pf + G dL G^T
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      E[i * ld + j] = GE[i * ld + j];
    }
    __syncwarp();
  }
  if (active) {
    double* e = elems + (bb * p.nchunk + c) * ps_smooth_elem(d);
    s2g<G>(e, E, ld, d, d);
    s2g<G>(e + d * d, Ls, ld, d, d);
    for (int i = gl; i < d; i += G) e[2 * d * d + i] = gv[i];
  }
}

// smoother scan (suffix): out[c] = in[c] o in[c + stride]   (identity on the right past the end)
//   E = E_i E_j ; g = E_i g_j + g_i ; L = E_i L_j E_i^T + L_i   (parallel_rts_smoother.py:39-55)
struct SsLayout { int d, ld, Ei, Li, Ej, Lj, W, Eo, gi, gj, bar, total; };
static SsLayout ss_layout(int d) {
  SsLayout L{};
  L.d = d; L.ld = (d % 2 == 0) ? d + 2 : (d | 1);
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  const int dd = d * L.ld;
  L.Ei = take(dd); L.Li = take(dd); L.Ej = take(dd); L.Lj = take(dd); L.W = take(dd); L.Eo = take(dd);
  L.gi = take(d); L.gj = take(d);
  L.bar = take(2);
  L.total = off;
  return L;
}

// result (Eo, gi, Lj) = (Ei, gi, Li) o (Ej, gj, Lj), operands in shared memory
template <int G>
__device__ __forceinline__ void smooth_combine(double* sm, const SsLayout& L) {
  const int d = L.d, ld = L.ld;
  double* Ei = sm + L.Ei; double* Li = sm + L.Li; double* Ej = sm + L.Ej; double* Lj = sm + L.Lj;
  double* W = sm + L.W; double* Eo = sm + L.Eo; double* gi = sm + L.gi; double* gj = sm + L.gj;
  mm<G, false, false>(W, ld, Ei, ld, Lj, ld, d, d, d, nullptr, 0, 1.0);       // Ei Lj
  mm<G, false, false>(Eo, ld, Ei, ld, Ej, ld, d, d, d, nullptr, 0, 1.0);      // Ei Ej
  mv<G, false>(gi, Ei, ld, gj, d, d, gi, 1.0);                                 // Ei gj + gi
  __syncwarp();
  mm<G, false, true>(Lj, ld, W, ld, Ei, ld, d, d, d, Li, ld, 1.0);             // Ei Lj Ei^T + Li
  __syncwarp();
}

template <int G>
__device__ __forceinline__ uint32_t load_smooth_elem(double* sm, const SsLayout& L, bool right,
                                                     const double* __restrict__ e) {
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  double* E = sm + (right ? L.Ej : L.Ei); double* Lm = sm + (right ? L.Lj : L.Li); double* g = sm + (right ? L.gj : L.gi);
  if ((d & 1) == 0) {
    if (gl == 0) {
      uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
      ptx::fence_proxy_async(ptx::space_shared);
      tma_rows(E, ld, e, d, d, bar);
      tma_rows(Lm, ld, e + d * d, d, d, bar);
      tma_rows(g, d, e + 2 * d * d, 1, d, bar);
    }
    return (uint32_t)((2 * d * d + d) * sizeof(double));
  }
  g2s<G>(E, ld, e, d, d);
  g2s<G>(Lm, ld, e + d * d, d, d);
  for (int i = gl; i < d; i += G) g[i] = e[2 * d * d + i];
  return 0;
}
// right operand = identity map (E = I, g = 0, L = 0) or a bare state (E = 0, g = m, L = P)
template <int G>
__device__ __forceinline__ void right_special(double* sm, const SsLayout& L, const double* __restrict__ mstate,
                                              const double* __restrict__ Pstate) {
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  for (int idx = gl; idx < d * d; idx += G) {
    const int i = idx / d, j = idx - i * d;
    sm[L.Ej + i * ld + j] = (!Pstate && i == j) ? 1.0 : 0.0;
    sm[L.Lj + i * ld + j] = Pstate ? Pstate[idx] : 0.0;
  }
  for (int i = gl; i < d; i += G) sm[L.gj + i] = mstate ? mstate[i] : 0.0;
}

// Hillis-Steele step of the suffix scan: out[c] = in[c] o in[c + stride]
template <int G>
__global__ void ps_smooth_scan_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t B,
                                      int64_t nchunk, int64_t stride, const SsLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = B * nchunk;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t b = g / nchunk, c = g % nchunk;
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int64_t ns = ps_smooth_elem(d);
  tma_bar_init<G>(reinterpret_cast<uint64_t*>(sm + L.bar));
  uint32_t parity = 0;
  uint32_t bytes = load_smooth_elem<G>(sm, L, false, in + (b * nchunk + c) * ns);
  if (c + stride < nchunk) bytes += load_smooth_elem<G>(sm, L, true, in + (b * nchunk + c + stride) * ns);
  else right_special<G>(sm, L, nullptr, nullptr);
  tma_finish<G>(sm, L.bar, bytes, parity);
  smooth_combine<G>(sm, L);
  if (!active) return;
  double* eo = out + (b * nchunk + c) * ns;
  s2g<G>(eo, sm + L.Eo, ld, d, d);
  for (int idx = gl; idx < d * d; idx += G) {
    const int i = idx / d, j = idx - i * d;
    eo[d * d + idx] = 0.5 * (sm[L.Lj + i * ld + j] + sm[L.Lj + j * ld + i]);
  }
  for (int i = gl; i < d; i += G) eo[2 * d * d + i] = sm[L.gi + i];
}

// apply: boundary of chunk c - 1 (= smoothed state at the first step of chunk c) = suffix[c] o start state,
// for c in [1, nchunk); the boundary of the last chunk is the start state itself.
// start state [B, d], [B, d, d]: the smoothed state one step past the end of this time range (the carried
// state of the next time shard, or the terminal (mf, Pf)[T - 1], whose step has dt = 0).
template <int G>
__global__ void ps_smooth_apply_kernel(const double* __restrict__ suffix, int64_t B, int64_t nchunk,
                                       const double* __restrict__ start_m, const double* __restrict__ start_P,
                                       double* __restrict__ bnd_m, double* __restrict__ bnd_P, const SsLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const int64_t n = B * nchunk;
  const bool active = gid < n;
  const int64_t g = active ? gid : n - 1;
  const int64_t b = g / nchunk, c = g % nchunk;
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int64_t ns = ps_smooth_elem(d);
  const int64_t cs = (c + 1 < nchunk) ? c + 1 : c;          // the last chunk's slot is written directly below
  tma_bar_init<G>(reinterpret_cast<uint64_t*>(sm + L.bar));
  uint32_t parity = 0;
  const uint32_t bytes = load_smooth_elem<G>(sm, L, false, suffix + (b * nchunk + cs) * ns);
  right_special<G>(sm, L, start_m + b * d, start_P + b * d * d);
  tma_finish<G>(sm, L.bar, bytes, parity);
  smooth_combine<G>(sm, L);
  if (!active) return;
  double* om = bnd_m + (b * nchunk + c) * d;
  double* oP = bnd_P + (b * nchunk + c) * d * d;
  if (c + 1 < nchunk) {
    for (int i = gl; i < d; i += G) om[i] = sm[L.gi + i];
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      oP[idx] = 0.5 * (sm[L.Lj + i * ld + j] + sm[L.Lj + j * ld + i]);
    }
  } else {
    for (int i = gl; i < d; i += G) om[i] = start_m[b * d + i];
    for (int idx = gl; idx < d * d; idx += G) oP[idx] = start_P[b * d * d + idx];
  }
}

// fold K shard totals onto a state, one group per series.
//   filter:   state <- state (x) totals[k], k = 0 .. K-1     (totals [K, B, ne])
template <int G>
__global__ void ps_filter_fold_kernel(const double* __restrict__ totals, int64_t B, int64_t K,
                                      const double* __restrict__ m0, int64_t m0_bs,
                                      const double* __restrict__ P0, int64_t P0_bs,
                                      double* __restrict__ m_out, double* __restrict__ P_out, const PcLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const bool active = gid < B;
  const int64_t b = active ? gid : B - 1;
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int64_t ne = ps_filter_elem(d);
  left_special<G>(sm, L, m0 + b * m0_bs, P0 + b * P0_bs);
  tma_bar_init<G>(reinterpret_cast<uint64_t*>(sm + L.bar));
  uint32_t parity = 0;
  for (int64_t k = 0; k < K; ++k) {
    const uint32_t bytes = load_filter_elem<G>(sm, L, true, totals + (k * B + b) * ne);
    tma_finish<G>(sm, L.bar, bytes, parity);
    filter_combine<G>(sm, L);
    // result (A = 0, b, C, J = 0, eta = 0) becomes the next left operand
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      sm[L.Ci + i * ld + j] = sm[L.Cj + i * ld + j];
    }
    for (int i = gl; i < d; i += G) sm[L.bi + i] = sm[L.bj + i];
    __syncwarp();
  }
  if (active) {
    for (int i = gl; i < d; i += G) m_out[b * d + i] = sm[L.bi + i];
    s2g<G>(P_out + b * d * d, sm + L.Ci, ld, d, d);
  }
}
//   smoother: state <- totals[k] o state, k = K-1 .. 0       (totals [K, B, ns], shard order in time)
template <int G>
__global__ void ps_smooth_fold_kernel(const double* __restrict__ totals, int64_t B, int64_t K,
                                      const double* __restrict__ m0, const double* __restrict__ P0,
                                      double* __restrict__ m_out, double* __restrict__ P_out, const SsLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int gpb = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const int64_t gid = (int64_t)blockIdx.x * gpb + g_in_block;
  const bool active = gid < B;
  const int64_t b = active ? gid : B - 1;
  const int d = L.d, ld = L.ld, gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int64_t ns = ps_smooth_elem(d);
  right_special<G>(sm, L, m0 + b * d, P0 + b * d * d);
  tma_bar_init<G>(reinterpret_cast<uint64_t*>(sm + L.bar));
  uint32_t parity = 0;
  for (int64_t k = K - 1; k >= 0; --k) {
    const uint32_t bytes = load_smooth_elem<G>(sm, L, false, totals + (k * B + b) * ns);
    tma_finish<G>(sm, L.bar, bytes, parity);
    smooth_combine<G>(sm, L);            // (Eo, gi, Lj); with Ej = 0 the map part stays 0
    for (int i = gl; i < d; i += G) sm[L.gj + i] = sm[L.gi + i];
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      sm[L.Ej + i * ld + j] = 0.0;
    }
    __syncwarp();
  }
  if (active) {
    for (int i = gl; i < d; i += G) m_out[b * d + i] = sm[L.gj + i];
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      P_out[b * d * d + idx] = 0.5 * (sm[L.Lj + i * ld + j] + sm[L.Lj + j * ld + i]);
    }
  }
}

// ------------------------------------------------------------------------------------ small kernels
// polish: boundary of chunk c (c >= 1) <- replayed filtered state at the last step of chunk c - 1
__global__ void ps_gather_bnd_kernel(const double* __restrict__ mf, const double* __restrict__ Pf, int64_t B,
                                     int64_t nchunk, int64_t chunk_len, int64_t sbs, int64_t sts, int d,
                                     double* __restrict__ bnd_m, double* __restrict__ bnd_P,
                                     const int* __restrict__ prev_changed = nullptr) {
  if (prev_changed && *prev_changed == 0) return;          // the previous fix-up pass changed nothing
  const int64_t n = B * nchunk * (int64_t)(d * d + d);
  const int per = d * d + d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / per;
    const int e = (int)(i - v * per);
    const int64_t b = v / nchunk, c = v % nchunk;
    if (c == 0) continue;
    const int64_t row = b * sbs + (c * chunk_len - 1) * sts;
    if (e < d) bnd_m[v * d + e] = mf[row * d + e];
    else bnd_P[v * d * d + (e - d)] = Pf[row * d * d + (e - d)];
  }
}

// smoother polish: boundary of chunk c (c < nchunk - 1) <- replayed smoothed state at the first step of chunk c + 1
__global__ void ps_gather_bnd_next_kernel(const double* __restrict__ ms, const double* __restrict__ Ps, int64_t B,
                                          int64_t nchunk, int64_t chunk_len, int64_t sbs, int64_t sts, int d,
                                          double* __restrict__ bnd_m, double* __restrict__ bnd_P) {
  const int64_t n = B * nchunk * (int64_t)(d * d + d);
  const int per = d * d + d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / per;
    const int e = (int)(i - v * per);
    const int64_t b = v / nchunk, c = v % nchunk;
    if (c == nchunk - 1) continue;
    const int64_t row = b * sbs + ((c + 1) * chunk_len) * sts;
    if (e < d) bnd_m[v * d + e] = ms[row * d + e];
    else bnd_P[v * d * d + (e - d)] = Ps[row * d * d + (e - d)];
  }
}

// lml[b] = sum_k lml_k[b, k], two deterministic stages: one 256-thread block per (series, segment of
// `seg` steps) -> partial[b, s]; then one warp per series over the partials.
__global__ void __launch_bounds__(256) ps_lml_partial_kernel(const double* __restrict__ lml_k, int64_t T, int64_t sbs,
                                                             int64_t sts, int64_t seg, int64_t nseg,
                                                             double* __restrict__ partial,
                                                             const double* __restrict__ sub = nullptr) {
  __shared__ double red[8];
  // flat grid of B * nseg blocks (gridDim.y is capped at 65,535: B >= 65,536 series must not fail)
  const int64_t b = (int64_t)blockIdx.x / nseg, sg = (int64_t)blockIdx.x % nseg;
  const int64_t k0 = sg * seg, k1 = (k0 + seg < T) ? k0 + seg : T;
  double s = 0.0;
  for (int64_t k = k0 + threadIdx.x; k < k1; k += 256) s += lml_k[b * sbs + k * sts] - (sub ? sub[b * sbs + k * sts] : 0.0);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[b * nseg + sg] = t;
  }
}
// the same partial sums for the time-major step layout (sbs == 1: the series index is the fastest one): a block takes
// 32 CONSECUTIVE series x one segment, every warp reads 32 series of one step -- 256 contiguous bytes -- where the kernel
// above reads one 8-byte word per 32-byte sector (ncu: 107 long-scoreboard stalls per issue, 1.9 TB/s of useful bytes).
__global__ void __launch_bounds__(256) ps_lml_partial_tm_kernel(const double* __restrict__ lml_k, int64_t B, int64_t T,
                                                                int64_t sts, int64_t seg, int64_t nseg,
                                                                double* __restrict__ partial,
                                                                const double* __restrict__ sub = nullptr) {
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t nb32 = (B + 31) / 32;
  const int64_t b = ((int64_t)blockIdx.x % nb32) * 32 + lane, sg = (int64_t)blockIdx.x / nb32;
  const int64_t k0 = sg * seg, k1 = (k0 + seg < T) ? k0 + seg : T;
  double s0 = 0.0, s1 = 0.0;
  if (b < B) {
    int64_t k = k0 + w;
    if (sub) {                   // sum of (x - sub): the ELBO's data ELL minus surrogate ELL
      for (; k + 8 < k1; k += 16) {
        s0 += lml_k[b + k * sts] - sub[b + k * sts];
        s1 += lml_k[b + (k + 8) * sts] - sub[b + (k + 8) * sts];
      }
      if (k < k1) s0 += lml_k[b + k * sts] - sub[b + k * sts];
    } else {
      for (; k + 8 < k1; k += 16) { s0 += lml_k[b + k * sts]; s1 += lml_k[b + (k + 8) * sts]; }
      if (k < k1) s0 += lml_k[b + k * sts];
    }
  }
  red[w][lane] = s0 + s1;
  __syncthreads();
  if (w == 0 && b < B) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += red[q][lane];
    partial[b * nseg + sg] = t;
  }
}
__global__ void ps_lml_final_kernel(const double* __restrict__ partial, int64_t B, int64_t nseg,
                                    double* __restrict__ lml) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= B) return;
  double s = 0.0;
  for (int64_t k = lane; k < nseg; k += 32) s += partial[w * nseg + k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) lml[w] = s;
}
static inline int64_t ps_lml_seg(int64_t T) { return T < 4096 ? T : 4096; }
static inline int64_t ps_lml_nseg(int64_t T) { const int64_t sg = ps_lml_seg(T); return (T + sg - 1) / sg; }
// lml[b] = sum_k lml_k[b, k].  Time-major layout with a free scratch of `big_n` doubles: 512-step segments summed by
// the coalesced kernel (enough blocks to fill the GPU), else 4096-step segments by the strided one into `small`.
static int ps_lml_sum(cudaStream_t st, const double* lml_k, int64_t B, int64_t T, int64_t sbs, int64_t sts,
                      double* small, double* big, int64_t big_n, double* lml, const double* sub = nullptr) {
  const int64_t seg_tm = 512, nseg_tm = (T + seg_tm - 1) / seg_tm;
  const double* partial;
  int64_t nseg;
  if (sbs == 1 && B >= 32 && big && B * nseg_tm <= big_n) {
    nseg = nseg_tm;
    partial = big;
    ps_lml_partial_tm_kernel<<<(unsigned)(nseg * ((B + 31) / 32)), 256, 0, st>>>(lml_k, B, T, sts, seg_tm, nseg, big, sub);
  } else {
    const int64_t seg = ps_lml_seg(T);
    nseg = ps_lml_nseg(T);
    partial = small;
    ps_lml_partial_kernel<<<(unsigned)(nseg * B), 256, 0, st>>>(lml_k, T, sbs, sts, seg, nseg, small, sub);
  }
  int rc = cuda_status(cudaGetLastError(), "ps_lml_partial_kernel launch");
  if (rc) return rc;
  const int64_t threads = B * 32;
  ps_lml_final_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(partial, B, nseg, lml);
  return cuda_status(cudaGetLastError(), "ps_lml_final_kernel launch");
}

// dst[r * dst_stride + i] = src[r * src_stride + i], i < n, r < rows   (src_stride may be 0 = broadcast)
__global__ void ps_copy_rows_kernel(double* __restrict__ dst, int64_t dst_stride, const double* __restrict__ src,
                                    int64_t src_stride, int64_t n, int64_t rows) {
  const int64_t total = n * rows;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n, c = i - r * n;
    dst[r * dst_stride + c] = src[r * src_stride + c];
  }
}
static int ps_copy_rows(cudaStream_t st, double* dst, int64_t dst_stride, const double* src, int64_t src_stride,
                        int64_t n, int64_t rows, const char* what) {
  const int64_t total = n * rows;
  if (total <= 0) return PHYSS_OK;
  const int64_t want = (total + 255) / 256;
  ps_copy_rows_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(dst, dst_stride, src, src_stride, n, rows);
  return cuda_status(cudaGetLastError(), what);
}

// --------------------------------------------------------------------------------------- launchers
template <typename K>
static int ps_launch_cfg(K kernel, size_t per_group_bytes, int G, int64_t ngroups, int& threads, size_t& smem,
                         int64_t& grid, const char* what) {
  threads = 128;
  while (threads > G && per_group_bytes * (threads / G) > 200 * 1024) threads /= 2;
  smem = per_group_bytes * (threads / G);
  if (smem > 200 * 1024)
    return set_error(PHYSS_ERR_UNSUPPORTED, "pscan: state dimension too large for the shared-memory path");
  const int gpb = threads / G;
  grid = (ngroups + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return cuda_status(e, what);
}

#define PS_LAUNCH(KERNEL, LAYOUT, NGROUPS, ...)                                                       \
  do {                                                                                                \
    int threads_; size_t smem_; int64_t grid_;                                                        \
    int rc_ = ps_launch_cfg(KERNEL, (size_t)(LAYOUT).total * sizeof(double), G, (NGROUPS), threads_,  \
                            smem_, grid_, "cudaFuncSetAttribute(" #KERNEL ")");                        \
    if (rc_) return rc_;                                                                              \
    KERNEL<<<(unsigned)grid_, threads_, smem_, st>>>(__VA_ARGS__);                                    \
    return cuda_status(cudaGetLastError(), #KERNEL " launch");                                        \
  } while (0)

template <int G, bool GIVEN>
static int run_filter_summary(cudaStream_t st, const SeqFilterArgs& a, const PsLayout& L, bool hid, int64_t cfirst,
                              int64_t ccount, double* elems) {
  PS_LAUNCH((ps_filter_summary_kernel<G, GIVEN>), L, a.B * ccount, a, L, hid, cfirst, ccount, elems);
}
template <int G>
static int run_filter_scan(cudaStream_t st, const double* in, double* out, int64_t B, int64_t nchunk, int64_t nsum,
                           int64_t stride, const PcLayout& L) {
  if (L.d == 8) PS_LAUNCH((ps_filter_scan_kernel<G, 8>), L, B * nsum, in, out, B, nchunk, nsum, stride, L);
  if (L.d == 6) PS_LAUNCH((ps_filter_scan_kernel<G, 6>), L, B * nsum, in, out, B, nchunk, nsum, stride, L);
  if (L.d == 12) PS_LAUNCH((ps_filter_scan_kernel<G, 12>), L, B * nsum, in, out, B, nchunk, nsum, stride, L);
  PS_LAUNCH((ps_filter_scan_kernel<G>), L, B * nsum, in, out, B, nchunk, nsum, stride, L);
}
// ---------------------------------------------------------------------------------------------------------
// Register forms of the two scan steps for d <= 3: one THREAD per element pair.  A d = 2 filter element is 16
// doubles; the lane-group kernels above spend their time on staging and group synchronisation there (measured
// 82 us per pass for 79k elements, 20 % of the CVI step of config 4), while the combine itself is a few dozen
// multiply-adds.  Same operators (parallel_kalman_filter.py:178-220 with the general solves of :201-211 and
// force_symmetric :216-219; parallel_rts_smoother.py:39-55).
template <int D>
__device__ __forceinline__ void small_solve(double (&M)[D][D], double (&X)[D][D]) {   // X <- M^-1 X (partial pivoting)
#pragma unroll
  for (int k = 0; k < D; ++k) {
    int pv = k;
    double best = fabs(M[k][k]);
#pragma unroll
    for (int r = k + 1; r < D; ++r) {
      const double v = fabs(M[r][k]);
      if (v > best) { best = v; pv = r; }
    }
#pragma unroll
    for (int r = k + 1; r < D; ++r) {
      const bool sw = (r == pv);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double a = M[k][c], b = M[r][c];
        M[k][c] = sw ? b : a;
        M[r][c] = sw ? a : b;
        const double xa = X[k][c], xb = X[r][c];
        X[k][c] = sw ? xb : xa;
        X[r][c] = sw ? xa : xb;
      }
    }
    const double inv = 1.0 / M[k][k];
#pragma unroll
    for (int r = k + 1; r < D; ++r) {
      const double f = M[r][k] * inv;
#pragma unroll
      for (int c = k + 1; c < D; ++c) M[r][c] = fma(-f, M[k][c], M[r][c]);
#pragma unroll
      for (int c = 0; c < D; ++c) X[r][c] = fma(-f, X[k][c], X[r][c]);
    }
  }
#pragma unroll
  for (int k = D - 1; k >= 0; --k) {
    const double inv = 1.0 / M[k][k];
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double t = X[k][c];
#pragma unroll
      for (int r = k + 1; r < D; ++r) t = fma(-M[k][r], X[r][c], t);
      X[k][c] = t * inv;
    }
  }
}

// o = l (x) r for two filtering elements (A, C, J, b, eta) stored as NE = 3 D^2 + 2 D consecutive doubles; l, r, o
// may live in global or shared memory, o may not alias l or r.
template <int D>
__device__ __forceinline__ void ps_filter_combine(const double* __restrict__ l, const double* __restrict__ r,
                                                  double* __restrict__ o) {
  double Ai[D][D], Ci[D][D], Ji[D][D], Aj[D][D], Cj[D][D], Jj[D][D], bi[D], ei[D], bj[D], ej[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      Ai[i][j] = l[i * D + j]; Ci[i][j] = l[D * D + i * D + j]; Ji[i][j] = l[2 * D * D + i * D + j];
      Aj[i][j] = r[i * D + j]; Cj[i][j] = r[D * D + i * D + j]; Jj[i][j] = r[2 * D * D + i * D + j];
    }
    bi[i] = l[3 * D * D + i]; ei[i] = l[3 * D * D + D + i];
    bj[i] = r[3 * D * D + i]; ej[i] = r[3 * D * D + D + i];
  }
  // M1 = I + Ci Jj ;  X1 = M1^-T Aj^T (= (Aj M1^-1)^T) ;  X2 = M1^-1 Ai
  double M1[D][D], M1T[D][D], X1[D][D], X2[D][D], t1[D], t2[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double a1 = bi[i], a2 = ej[i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) acc = fma(Ci[i][q], Jj[q][j], acc);
      M1[i][j] = acc;
      M1T[j][i] = acc;
      X1[i][j] = Aj[j][i];
      X2[i][j] = Ai[i][j];
      a1 = fma(Ci[i][j], ej[j], a1);                        // t1 = bi + Ci ej
      a2 = fma(-Jj[i][j], bi[j], a2);                       // t2 = ej - Jj bi
    }
    t1[i] = a1;
    t2[i] = a2;
  }
  small_solve<D>(M1T, X1);
  small_solve<D>(M1, X2);
  // A_out = X1^T Ai ; W = X1^T Ci ; C_out = W Aj^T + Cj ; b_out = X1^T t1 + bj
  // V = X2^T Jj ; J_out = V Ai + Ji ; eta_out = X2^T t2 + ei
  double W[D][D], V[D][D], Co[D][D], Jo[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double a1 = bj[i], a2 = ei[i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double ao = 0.0, w = 0.0, v = 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) {
        ao = fma(X1[q][i], Ai[q][j], ao);
        w = fma(X1[q][i], Ci[q][j], w);
        v = fma(X2[q][i], Jj[q][j], v);
      }
      o[i * D + j] = ao;
      W[i][j] = w;
      V[i][j] = v;
      a1 = fma(X1[j][i], t1[j], a1);
      a2 = fma(X2[j][i], t2[j], a2);
    }
    o[3 * D * D + i] = a1;
    o[3 * D * D + D + i] = a2;
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double co = Cj[i][j], jo = Ji[i][j];
#pragma unroll
      for (int q = 0; q < D; ++q) {
        co = fma(W[i][q], Aj[j][q], co);
        jo = fma(V[i][q], Ai[q][j], jo);
      }
      Co[i][j] = co;
      Jo[i][j] = jo;
    }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      o[D * D + i * D + j] = 0.5 * (Co[i][j] + Co[j][i]);
      o[2 * D * D + i * D + j] = 0.5 * (Jo[i][j] + Jo[j][i]);
    }
}

template <int D>
__global__ void __launch_bounds__(128) ps_filter_scan_reg_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                                 int64_t B, int64_t nchunk, int64_t nsum, int64_t stride) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * nsum) return;
  const int64_t b = gid / nsum, c = gid % nsum;
  constexpr int NE = 3 * D * D + 2 * D;
  const double* r = in + (b * nchunk + c) * NE;
  double* o = out + (b * nchunk + c) * NE;
  if (c < stride) {                                       // identity on the left
#pragma unroll
    for (int i = 0; i < NE; ++i) o[i] = r[i];
    return;
  }
  const double* l = in + (b * nchunk + c - stride) * NE;
  ps_filter_combine<D>(l, r, o);
}

// suffix scan step of the smoother: out[c] = in[c] o in[c + stride]  (identity on the right past the end)
// o = l (x) r for two smoothing elements (E, L, g), NS = 2 D^2 + D consecutive doubles each
template <int D>
__device__ __forceinline__ void ps_smooth_combine(const double* __restrict__ l, const double* __restrict__ r,
                                                  double* __restrict__ o) {
  double Ei[D][D], Li[D][D], Ej[D][D], Lj[D][D], gi[D], gj[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      Ei[i][j] = l[i * D + j]; Li[i][j] = l[D * D + i * D + j];
      Ej[i][j] = r[i * D + j]; Lj[i][j] = r[D * D + i * D + j];
    }
    gi[i] = l[2 * D * D + i];
    gj[i] = r[2 * D * D + i];
  }
  double W[D][D], Lo[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double g = gi[i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double w = 0.0, eo = 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) {
        w = fma(Ei[i][q], Lj[q][j], w);                    // Ei Lj
        eo = fma(Ei[i][q], Ej[q][j], eo);                  // Ei Ej
      }
      W[i][j] = w;
      o[i * D + j] = eo;
      g = fma(Ei[i][j], gj[j], g);                          // Ei gj + gi
    }
    o[2 * D * D + i] = g;
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double acc = Li[i][j];
#pragma unroll
      for (int q = 0; q < D; ++q) acc = fma(W[i][q], Ei[j][q], acc);   // Ei Lj Ei^T + Li
      Lo[i][j] = acc;
    }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) o[D * D + i * D + j] = 0.5 * (Lo[i][j] + Lo[j][i]);
}

template <int D>
__global__ void __launch_bounds__(128) ps_smooth_scan_reg_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                                 int64_t B, int64_t nchunk, int64_t stride) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * nchunk) return;
  const int64_t b = gid / nchunk, c = gid % nchunk;
  constexpr int NS = 2 * D * D + D;
  const double* l = in + (b * nchunk + c) * NS;
  double* o = out + (b * nchunk + c) * NS;
  if (c + stride >= nchunk) {
#pragma unroll
    for (int i = 0; i < NS; ++i) o[i] = l[i];
    return;
  }
  const double* r = in + (b * nchunk + c + stride) * NS;
  ps_smooth_combine<D>(l, r, o);
}

// All Hillis-Steele passes of one series in ONE launch: a block per series, its <= 128 chunk elements ping-pong between
// two shared-memory buffers (element stride padded to an odd number of doubles: conflict-free), one __syncthreads per
// pass.  Replaces log2(nchunk) launches that each round-trip every element through L2 (7 x 20 us -> one ~25 us kernel at
// 1000 series x 79 chunks, d = 2).  dst may alias in (every element is read before any is written).
template <int D, bool SUFFIX>
__global__ void __launch_bounds__(128) ps_scan_fused_kernel(const double* in, double* dst, int64_t nchunk, int n) {
  constexpr int NE = SUFFIX ? 2 * D * D + D : 3 * D * D + 2 * D;
  constexpr int LDE = NE | 1;
  extern __shared__ __align__(16) double ps_fused_sm[];
  double* cur = ps_fused_sm;
  double* nxt = ps_fused_sm + 128 * LDE;
  const int c = threadIdx.x;
  const double* src = in + (int64_t)blockIdx.x * nchunk * NE;
  for (int idx = threadIdx.x; idx < n * NE; idx += 128) cur[(idx / NE) * LDE + idx % NE] = src[idx];
  __syncthreads();
  for (int stride = 1; stride < n; stride *= 2) {
    if (c < n) {
      double* o = nxt + c * LDE;
      const bool edge = SUFFIX ? (c + stride >= n) : (c < stride);
      if (edge) {
#pragma unroll
        for (int i = 0; i < NE; ++i) o[i] = cur[c * LDE + i];
      } else if (SUFFIX) {
        ps_smooth_combine<D>(cur + c * LDE, cur + (c + stride) * LDE, o);
      } else {
        ps_filter_combine<D>(cur + (c - stride) * LDE, cur + c * LDE, o);
      }
    }
    __syncthreads();
    double* t = cur; cur = nxt; nxt = t;
  }
  double* out = dst + (int64_t)blockIdx.x * nchunk * NE;
  for (int idx = threadIdx.x; idx < n * NE; idx += 128) out[idx] = cur[(idx / NE) * LDE + idx % NE];
}

// apply steps in registers (same role as ps_filter_apply_kernel / ps_smooth_apply_kernel above)
template <int D>
__global__ void __launch_bounds__(128) ps_filter_apply_reg_kernel(const double* __restrict__ prefix, int64_t B,
                                                                  int64_t nchunk, int64_t nsum,
                                                                  const double* __restrict__ m0, int64_t m0_bs,
                                                                  const double* __restrict__ P0, int64_t P0_bs,
                                                                  double* __restrict__ bnd_m, double* __restrict__ bnd_P) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * nsum) return;
  const int64_t b = gid / nsum, c = gid % nsum;
  constexpr int NE = 3 * D * D + 2 * D;
  const double* r = prefix + (b * nchunk + c) * NE;
  double Ci[D][D], bi[D], Aj[D][D], Cj[D][D], Jj[D][D], bj[D], ej[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      Ci[i][j] = P0[b * P0_bs + i * D + j];
      Aj[i][j] = r[i * D + j]; Cj[i][j] = r[D * D + i * D + j]; Jj[i][j] = r[2 * D * D + i * D + j];
    }
    bi[i] = m0[b * m0_bs + i];
    bj[i] = r[3 * D * D + i]; ej[i] = r[3 * D * D + D + i];
  }
  // left = bare state (A = 0, b = m0, C = P0, J = 0, eta = 0):  M1 = I + P0 Jj ; X1 = M1^-T Aj^T ; t1 = m0 + P0 ej
  double M1T[D][D], X1[D][D], t1[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double a1 = bi[i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) acc = fma(Ci[i][q], Jj[q][j], acc);
      M1T[j][i] = acc;
      X1[i][j] = Aj[j][i];
      a1 = fma(Ci[i][j], ej[j], a1);
    }
    t1[i] = a1;
  }
  small_solve<D>(M1T, X1);
  double W[D][D], Co[D][D];
  double* om = bnd_m + (b * nchunk + c + 1) * D;
  double* oP = bnd_P + (b * nchunk + c + 1) * D * D;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double a1 = bj[i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double w = 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) w = fma(X1[q][i], Ci[q][j], w);
      W[i][j] = w;
      a1 = fma(X1[j][i], t1[j], a1);
    }
    om[i] = a1;                                            // b_out = X1^T t1 + bj
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double co = Cj[i][j];
#pragma unroll
      for (int q = 0; q < D; ++q) co = fma(W[i][q], Aj[j][q], co);
      Co[i][j] = co;                                       // C_out = W Aj^T + Cj
    }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) oP[i * D + j] = 0.5 * (Co[i][j] + Co[j][i]);
  if (c == 0) {                                            // boundary 0 is the start state itself
#pragma unroll
    for (int i = 0; i < D; ++i) {
      bnd_m[(b * nchunk) * D + i] = bi[i];
#pragma unroll
      for (int j = 0; j < D; ++j) bnd_P[(b * nchunk) * D * D + i * D + j] = Ci[i][j];
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128) ps_smooth_apply_reg_kernel(const double* __restrict__ suffix, int64_t B,
                                                                  int64_t nchunk, const double* __restrict__ start_m,
                                                                  const double* __restrict__ start_P,
                                                                  double* __restrict__ bnd_m, double* __restrict__ bnd_P) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * nchunk) return;
  const int64_t b = gid / nchunk, c = gid % nchunk;
  constexpr int NS = 2 * D * D + D;
  double* om = bnd_m + (b * nchunk + c) * D;
  double* oP = bnd_P + (b * nchunk + c) * D * D;
  double m[D], P[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    m[i] = start_m[b * D + i];
#pragma unroll
    for (int j = 0; j < D; ++j) P[i][j] = start_P[b * D * D + i * D + j];
  }
  if (c + 1 >= nchunk) {                                   // the last chunk starts from the start state itself
#pragma unroll
    for (int i = 0; i < D; ++i) {
      om[i] = m[i];
#pragma unroll
      for (int j = 0; j < D; ++j) oP[i * D + j] = P[i][j];
    }
    return;
  }
  const double* l = suffix + (b * nchunk + c + 1) * NS;    // (E, L, g) of chunks c + 1 .. end
  double E[D][D], W[D][D], Lo[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) E[i][j] = l[i * D + j];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double g = l[2 * D * D + i];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double w = 0.0;
#pragma unroll
      for (int q = 0; q < D; ++q) w = fma(E[i][q], P[q][j], w);
      W[i][j] = w;
      g = fma(E[i][j], m[j], g);
    }
    om[i] = g;                                             // E m + g
  }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double acc = l[D * D + i * D + j];
#pragma unroll
      for (int q = 0; q < D; ++q) acc = fma(W[i][q], E[j][q], acc);
      Lo[i][j] = acc;                                      // E P E^T + L
    }
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) oP[i * D + j] = 0.5 * (Lo[i][j] + Lo[j][i]);
}

static int run_filter_apply_reg(cudaStream_t st, int d, const double* prefix, int64_t B, int64_t nchunk, int64_t nsum,
                                const double* m0, int64_t m0_bs, const double* P0, int64_t P0_bs, double* bnd_m,
                                double* bnd_P) {
  const int64_t n = B * nsum;
  const unsigned grid = (unsigned)((n + 127) / 128);
#define PHYSS_AP(D_) ps_filter_apply_reg_kernel<D_><<<grid, 128, 0, st>>>(prefix, B, nchunk, nsum, m0, m0_bs, P0, P0_bs, bnd_m, bnd_P)
  if (d == 1) PHYSS_AP(1); else if (d == 2) PHYSS_AP(2); else if (d == 3) PHYSS_AP(3); else PHYSS_AP(4);
#undef PHYSS_AP
  return cuda_status(cudaGetLastError(), "ps_filter_apply_reg_kernel launch");
}
static int run_smooth_apply_reg(cudaStream_t st, int d, const double* suffix, int64_t B, int64_t nchunk,
                                const double* start_m, const double* start_P, double* bnd_m, double* bnd_P) {
  const int64_t n = B * nchunk;
  const unsigned grid = (unsigned)((n + 127) / 128);
#define PHYSS_AP(D_) ps_smooth_apply_reg_kernel<D_><<<grid, 128, 0, st>>>(suffix, B, nchunk, start_m, start_P, bnd_m, bnd_P)
  if (d == 1) PHYSS_AP(1); else if (d == 2) PHYSS_AP(2); else if (d == 3) PHYSS_AP(3); else PHYSS_AP(4);
#undef PHYSS_AP
  return cuda_status(cudaGetLastError(), "ps_smooth_apply_reg_kernel launch");
}

static bool ps_reg_scan(int d) { return d >= 1 && d <= 4 && !ps_force_grp(); }

static int run_filter_scan_reg(cudaStream_t st, int d, const double* in, double* out, int64_t B, int64_t nchunk,
                               int64_t nsum, int64_t stride) {
  const int64_t n = B * nsum;
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (d == 1) ps_filter_scan_reg_kernel<1><<<grid, 128, 0, st>>>(in, out, B, nchunk, nsum, stride);
  else if (d == 2) ps_filter_scan_reg_kernel<2><<<grid, 128, 0, st>>>(in, out, B, nchunk, nsum, stride);
  else if (d == 3) ps_filter_scan_reg_kernel<3><<<grid, 128, 0, st>>>(in, out, B, nchunk, nsum, stride);
  else ps_filter_scan_reg_kernel<4><<<grid, 128, 0, st>>>(in, out, B, nchunk, nsum, stride);
  return cuda_status(cudaGetLastError(), "ps_filter_scan_reg_kernel launch");
}
static int run_smooth_scan_reg(cudaStream_t st, int d, const double* in, double* out, int64_t B, int64_t nchunk,
                               int64_t stride) {
  const int64_t n = B * nchunk;
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (d == 1) ps_smooth_scan_reg_kernel<1><<<grid, 128, 0, st>>>(in, out, B, nchunk, stride);
  else if (d == 2) ps_smooth_scan_reg_kernel<2><<<grid, 128, 0, st>>>(in, out, B, nchunk, stride);
  else if (d == 3) ps_smooth_scan_reg_kernel<3><<<grid, 128, 0, st>>>(in, out, B, nchunk, stride);
  else ps_smooth_scan_reg_kernel<4><<<grid, 128, 0, st>>>(in, out, B, nchunk, stride);
  return cuda_status(cudaGetLastError(), "ps_smooth_scan_reg_kernel launch");
}

// fused scan (ps_scan_fused_kernel): all passes in one launch when a series has at most 128 elements to scan
static bool ps_fused_scan(int d, int64_t n) {
  static const bool off = getenv("PHYSS_PSCAN_NOFUSE") != nullptr;
  return ps_reg_scan(d) && n >= 2 && n <= 128 && !off;
}
template <int D, bool SUFFIX>
static int launch_scan_fused(cudaStream_t st, const double* in, double* dst, int64_t B, int64_t nchunk, int64_t n) {
  constexpr int NE = SUFFIX ? 2 * D * D + D : 3 * D * D + 2 * D;
  const size_t smem = 2 * 128 * (size_t)(NE | 1) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ps_scan_fused_kernel<D, SUFFIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "ps_scan_fused_kernel: shared memory attribute");
    configured = true;
  }
  ps_scan_fused_kernel<D, SUFFIX><<<(unsigned)B, 128, smem, st>>>(in, dst, nchunk, (int)n);
  return cuda_status(cudaGetLastError(), "ps_scan_fused_kernel launch");
}
template <bool SUFFIX>
static int run_scan_fused(cudaStream_t st, int d, const double* in, double* dst, int64_t B, int64_t nchunk, int64_t n) {
  if (d == 1) return launch_scan_fused<1, SUFFIX>(st, in, dst, B, nchunk, n);
  if (d == 2) return launch_scan_fused<2, SUFFIX>(st, in, dst, B, nchunk, n);
  if (d == 3) return launch_scan_fused<3, SUFFIX>(st, in, dst, B, nchunk, n);
  return launch_scan_fused<4, SUFFIX>(st, in, dst, B, nchunk, n);
}

template <int G>
static int run_filter_apply(cudaStream_t st, const double* prefix, int64_t B, int64_t nchunk, int64_t nsum,
                            const double* m0, int64_t m0_bs, const double* P0, int64_t P0_bs, double* bnd_m,
                            double* bnd_P, const PcLayout& L) {
  PS_LAUNCH((ps_filter_apply_kernel<G>), L, B * nsum, prefix, B, nchunk, nsum, m0, m0_bs, P0, P0_bs, bnd_m, bnd_P, L);
}
template <int G>
static int run_filter_fold(cudaStream_t st, const double* totals, int64_t B, int64_t K, const double* m0,
                           int64_t m0_bs, const double* P0, int64_t P0_bs, double* m_out, double* P_out,
                           const PcLayout& L) {
  PS_LAUNCH((ps_filter_fold_kernel<G>), L, B, totals, B, K, m0, m0_bs, P0, P0_bs, m_out, P_out, L);
}
template <int G, bool GIVEN>
static int run_smooth_summary(cudaStream_t st, const SeqSmoothArgs& a, const PsLayout& L, double* elems) {
  PS_LAUNCH((ps_smooth_summary_kernel<G, GIVEN>), L, a.B * a.chunk_count, a, L, elems);
}
template <int G>
static int run_smooth_scan(cudaStream_t st, const double* in, double* out, int64_t B, int64_t nchunk, int64_t stride,
                           const SsLayout& L) {
  PS_LAUNCH((ps_smooth_scan_kernel<G>), L, B * nchunk, in, out, B, nchunk, stride, L);
}
template <int G>
static int run_smooth_apply(cudaStream_t st, const double* suffix, int64_t B, int64_t nchunk, const double* sm_,
                            const double* sP, double* bnd_m, double* bnd_P, const SsLayout& L) {
  PS_LAUNCH((ps_smooth_apply_kernel<G>), L, B * nchunk, suffix, B, nchunk, sm_, sP, bnd_m, bnd_P, L);
}
template <int G>
static int run_smooth_fold(cudaStream_t st, const double* totals, int64_t B, int64_t K, const double* m0,
                           const double* P0, double* m_out, double* P_out, const SsLayout& L) {
  PS_LAUNCH((ps_smooth_fold_kernel<G>), L, B, totals, B, K, m0, P0, m_out, P_out, L);
}

#define PS_BY_G(FN, ...) (G == 8 ? FN<8>(__VA_ARGS__) : (G == 16 ? FN<16>(__VA_ARGS__) : FN<32>(__VA_ARGS__)))
#define PS_BY_G_GIVEN(FN, ...)                                                                            \
  (given ? (G == 8 ? FN<8, true>(__VA_ARGS__) : (G == 16 ? FN<16, true>(__VA_ARGS__) : FN<32, true>(__VA_ARGS__))) \
         : (G == 8 ? FN<8, false>(__VA_ARGS__) : (G == 16 ? FN<16, false>(__VA_ARGS__) : FN<32, false>(__VA_ARGS__))))

// ------------------------------------------------------------------------------------ workspace map
// [ e0 | e1 | bnd_m | bnd_P | start (d + d*d per series) | flag (2 doubles) | lml_k (optional) ]
struct PsWorkspace {
  double* e0; double* e1;
  double* bnd_m; double* bnd_P;
  double* start_m; double* start_P;
  int* flag;
  double* lml_partial;
  double* lml_k;
};
static int64_t ps_nchunk(int64_t T, int64_t chunk_len) { return (T + chunk_len - 1) / chunk_len; }

int64_t pscan_workspace_doubles(int64_t B, int64_t T, int d, int64_t chunk_len) {
  const int64_t nchunk = ps_nchunk(T, chunk_len);
  const int64_t sd = (int64_t)d + (int64_t)d * d;
  return 2 * B * nchunk * ps_filter_elem(d) + B * nchunk * sd + B * sd + 2 + B * ((T + 4095) / 4096 + 1) + B * T;
}
static PsWorkspace ps_carve(double* ws, int64_t B, int64_t T, int d, int64_t chunk_len) {
  const int64_t nchunk = ps_nchunk(T, chunk_len);
  PsWorkspace w{};
  double* p = ws;
  w.e0 = p; p += B * nchunk * ps_filter_elem(d);
  w.e1 = p; p += B * nchunk * ps_filter_elem(d);
  w.bnd_m = p; p += B * nchunk * d;
  w.bnd_P = p; p += B * nchunk * (int64_t)d * d;
  w.start_m = p; p += B * d;
  w.start_P = p; p += B * (int64_t)d * d;
  w.flag = reinterpret_cast<int*>(p); p += 2;
  w.lml_partial = p; p += B * ((T + 4095) / 4096 + 1);
  w.lml_k = p;
  return w;
}
// buffer that holds the scan result after the Hillis-Steele passes over n elements
static double* ps_scan_result(const PsWorkspace& w, int64_t n) {
  int passes = 0;
  for (int64_t stride = 1; stride < n; stride *= 2) ++passes;
  return (passes % 2 == 0) ? w.e0 : w.e1;
}

static int check_matern(int d, int disc_mode, int nblk) {
  if (disc_mode == PHYSS_DISC_GIVEN) return PHYSS_OK;
  if (disc_mode != PHYSS_DISC_MATERN)
    return set_error(PHYSS_ERR_UNSUPPORTED, "parallel-in-time forms: DISC_GIVEN or DISC_MATERN only");
  const int s = (nblk > 0) ? d / nblk : 0;
  if (s < 1 || s > 4 || s * nblk != d)
    return set_error(PHYSS_ERR_UNSUPPORTED, "DISC_MATERN needs equal blocks of size 1..4");
  return PHYSS_OK;
}

// -------------------------------------------------------------------------------------- filter entry
// local: chunk summaries + prefix scan (left in the workspace); with `total_out` also the element of the
// whole time range [B, 3 d^2 + 2 d] (multi-GPU: the summary this time shard contributes to the all-gather).
int pscan_filter_local(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, SeqFilterArgs a,
                       int64_t chunk_len, double* ws, double* total_out) {
  int rc = check_matern(d, disc_mode, nblk);
  if (rc) return rc;
  if (m > d) return set_error(PHYSS_ERR_UNSUPPORTED, "pscan filter: m > d is not supported");
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  const int64_t nchunk = ps_nchunk(a.T, chunk_len);
  const int64_t nsum = total_out ? nchunk : nchunk - 1;   // the last chunk's summary only feeds the total
  if (nsum <= 0) return PHYSS_OK;
  PsWorkspace w = ps_carve(ws, a.B, a.T, d, chunk_len);
  a.nchunk = nchunk; a.chunk_len = chunk_len;
  const int G = ps_group_size(d);
  const PsLayout L = ps_layout(d, m, given ? 0 : nblk, given, false);
  const PcLayout Lc = pc_layout(d);
  const int64_t nfull = a.T / chunk_len;
  const int64_t nfull_sum = nfull < nsum ? nfull : nsum;
  const bool reg = d <= 4 && seq_supported(d, m, disc_mode, nblk);     // register-resident summaries for d <= 4
  auto summarise = [&](int64_t first, int64_t count) -> int {
    if (count <= 0) return PHYSS_OK;
    if (reg) {
      SeqFilterArgs r = a;
      r.chunk_first = first; r.chunk_count = count;
      return seq_filter_summary(st, d, m, disc_mode, nblk, hid, r, w.e0);
    }
    if (rt_supported(d, m) && !ps_force_grp()) return rt_filter_summary(st, d, m, disc_mode, nblk, hid, a, first, count, w.e0);
    return PS_BY_G_GIVEN(run_filter_summary, st, a, L, hid, first, count, w.e0);
  };
  rc = summarise(0, nfull_sum);
  if (rc) return rc;
  rc = summarise(nfull_sum, nsum - nfull_sum);
  if (rc) return rc;
  double* in = w.e0; double* out = w.e1;
  const bool fused = ps_fused_scan(d, nsum) && a.B <= 0x7fffffff;
  if (fused) {                  // the result lands where the pass-by-pass ping-pong would have left it
    for (int64_t stride = 1; stride < nsum; stride *= 2) { double* t = in; in = out; out = t; }
    rc = run_scan_fused<false>(st, d, w.e0, in, a.B, nchunk, nsum);
    if (rc) return rc;
  }
  // few elements (one long series): a pass is as long as ONE combine, so every combine gets a whole warp
  static const int scan_g = [] { const char* e = getenv("PHYSS_PSCAN_SCAN_G"); return e ? atoi(e) : 0; }();
  const int Gscan = (scan_g == 8 || scan_g == 16 || scan_g == 32) ? scan_g : (a.B * nsum <= 148 * 64 ? 32 : G);
  for (int64_t stride = 1; !fused && stride < nsum; stride *= 2) {
    if (ps_reg_scan(d)) {
      rc = run_filter_scan_reg(st, d, in, out, a.B, nchunk, nsum, stride);
    } else {
      const int G = Gscan;      // (shadows the group size of the summary kernels inside PS_BY_G)
      rc = PS_BY_G(run_filter_scan, st, in, out, a.B, nchunk, nsum, stride, Lc);
    }
    if (rc) return rc;
    double* t = in; in = out; out = t;
  }
  if (total_out) {
    const int64_t ne = ps_filter_elem(d);
    rc = ps_copy_rows(st, total_out, ne, in + (nchunk - 1) * ne, nchunk * ne, ne, a.B, "pscan filter: copy of the range total");
    if (rc) return rc;
  }
  return PHYSS_OK;
}

// finish: boundaries from the start state, concurrent replay of all chunks, polish passes, lml.
// `had_total` must repeat whether pscan_filter_local was called with total_out (it decides which
// workspace buffer holds the prefixes).  start_m / start_P [B, d] / [B, d, d] or NULL = (m0, P0).
int pscan_filter_finish(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, SeqFilterArgs a,
                        int64_t chunk_len, double* ws, bool had_total, const double* start_m,
                        const double* start_P, int polish, double delta, int patience, int* status_out) {
  int rc = check_matern(d, disc_mode, nblk);
  if (rc) return rc;
  const int64_t nchunk = ps_nchunk(a.T, chunk_len);
  PsWorkspace w = ps_carve(ws, a.B, a.T, d, chunk_len);
  const int G = ps_group_size(d);
  const PcLayout Lc = pc_layout(d);
  const double* sm0 = start_m ? start_m : a.m0;
  const int64_t sm0_bs = start_m ? d : a.m0_bs;
  const double* sP0 = start_P ? start_P : a.P0;
  const int64_t sP0_bs = start_P ? (int64_t)d * d : a.P0_bs;
  cudaError_t e;
  if (nchunk > 1) {
    const double* prefix = ps_scan_result(w, had_total ? nchunk : nchunk - 1);
    rc = ps_reg_scan(d) ? run_filter_apply_reg(st, d, prefix, a.B, nchunk, nchunk - 1, sm0, sm0_bs, sP0, sP0_bs, w.bnd_m, w.bnd_P)
                        : PS_BY_G(run_filter_apply, st, prefix, a.B, nchunk, nchunk - 1, sm0, sm0_bs, sP0, sP0_bs, w.bnd_m, w.bnd_P, Lc);
    if (rc) return rc;
  } else {
    rc = ps_copy_rows(st, w.bnd_m, d, sm0, sm0_bs, d, a.B, "pscan filter: copy of the start mean");
    if (rc) return rc;
    rc = ps_copy_rows(st, w.bnd_P, (int64_t)d * d, sP0, sP0_bs, (int64_t)d * d, a.B, "pscan filter: copy of the start covariance");
    if (rc) return rc;
  }
  e = cudaMemsetAsync(w.flag, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return cuda_status(e, "pscan filter: flag reset");
  if (!a.lml_k) a.lml_k = w.lml_k;
  a.nchunk = nchunk; a.chunk_len = chunk_len;
  a.bnd_m = w.bnd_m; a.bnd_P = w.bnd_P; a.unconverged = w.flag;
  a.from_bnd = 1; a.fixup = 0; a.delta = delta; a.patience = patience;
  const int64_t nfull = a.T / chunk_len;
  auto replay = [&](int64_t first, int64_t count) -> int {
    if (count <= 0) return PHYSS_OK;
    SeqFilterArgs r = a;
    r.chunk_first = first; r.chunk_count = count;
    return run_filter_any(st, d, m, disc_mode, nblk, hid, r);
  };
  rc = replay(0, nfull);
  if (rc) return rc;
  rc = replay(nfull, nchunk - nfull);
  if (rc) return rc;
  // fix-up passes after a pass that changed nothing are no-ops: the register kernels flag a pass in which any recomputed
  // step disagreed (w.flag[1 + parity]); the next pass and its boundary gather return at once when that flag is clear
  const bool early_out = run_filter_is_seq(d, m, disc_mode, nblk, a.B, nchunk) && !getenv("PHYSS_PSCAN_NO_EARLY_OUT");
  for (int it = 0; it < polish && nchunk > 1; ++it) {
    const int64_t total = a.B * nchunk * (int64_t)(d * d + d);
    const int64_t want = (total + 255) / 256;
    const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
    int* cur = early_out ? w.flag + 1 + (it & 1) : nullptr;
    const int* prev = (early_out && it > 0) ? w.flag + 1 + ((it + 1) & 1) : nullptr;
    ps_gather_bnd_kernel<<<blocks, 256, 0, st>>>(a.mf, a.Pf, a.B, nchunk, chunk_len, a.sbs, a.sts, d, w.bnd_m, w.bnd_P, prev);
    rc = cuda_status(cudaGetLastError(), "ps_gather_bnd_kernel launch");
    if (rc) return rc;
    e = cudaMemsetAsync(w.flag, 0, sizeof(int), st);
    if (e == cudaSuccess && cur) e = cudaMemsetAsync(cur, 0, sizeof(int), st);
    if (e != cudaSuccess) return cuda_status(e, "pscan filter: flag reset");
    a.fixup = 1; a.pass_changed = cur; a.prev_changed = prev;
    rc = replay(1, nfull - 1);
    if (rc) return rc;
    if (nfull >= 1) { rc = replay(nfull, nchunk - nfull); if (rc) return rc; }
    a.fixup = 0; a.pass_changed = nullptr; a.prev_changed = nullptr;
  }
  {
    // scratch for the partial sums: the scan buffer that does NOT hold the prefixes (both are free once the boundaries
    // exist, but the caller of a time-sharded pass may still read the prefixes)
    double* other = (nchunk > 1 && ps_scan_result(w, had_total ? nchunk : nchunk - 1) == w.e1) ? w.e0 : w.e1;
    rc = ps_lml_sum(st, a.lml_k, a.B, a.T, a.sbs, a.sts, w.lml_partial, other, a.B * nchunk * ps_filter_elem(d), a.lml);
    if (rc) return rc;
  }
  if (status_out) {
    e = cudaMemcpyAsync(status_out, w.flag, sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_status(e, "pscan filter: status copy");
  }
  return PHYSS_OK;
}

// per-series sums over the time axis of a [B, T] array in either step layout (the ELBO's ELL sums); scratch:
// B * ceil(T / 512) doubles
int sum_steps(cudaStream_t st, int64_t B, int64_t T, int64_t sbs, int64_t sts, const double* x, const double* sub,
              double* scratch, double* out) {
  if (B <= 0 || T <= 0) return PHYSS_OK;
  return ps_lml_sum(st, x, B, T, sbs, sts, scratch, scratch, B * ((T + 511) / 512), out, sub);
}

// ---------------------------------------------------------------------------------- speculative mode
// No scan: every chunk c > 0 starts `warm` steps early from the initial state (m0, P0) -- for the reference's
// filters the stationary prior -- and discards those steps: the filter forgets its start at the rate the
// polish passes rely on anyway.  The same fix-up passes then (i) verify every chunk against a restart from
// the previous chunk's end state and (ii) repair it where the warm-up was too short; *status reports chunks
// that still disagreed in the last pass.  Costs (1 + warm / chunk_len) replays instead of summary + scan +
// replay; the caller falls back to the exact scan when the flag is raised.
int pscan_filter_spec(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, SeqFilterArgs a,
                      int64_t chunk_len, int64_t warm, int polish, double delta, int patience, double* ws,
                      int* status_out) {
  int rc = check_matern(d, disc_mode, nblk);
  if (rc) return rc;
  if (d > 32) return set_error(PHYSS_ERR_UNSUPPORTED, "pscan speculative mode: d <= 32 only");
  if (warm < 1 || warm > chunk_len) return set_error(PHYSS_ERR_BAD_ARG, "pscan speculative mode: need 1 <= warm <= chunk_len");
  if (polish < 1) polish = 1;                               // at least the verification pass
  const int64_t nchunk = ps_nchunk(a.T, chunk_len);
  PsWorkspace w = ps_carve(ws, a.B, a.T, d, chunk_len);
  cudaError_t e = cudaMemsetAsync(w.flag, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return cuda_status(e, "pscan filter: flag reset");
  if (!a.lml_k) a.lml_k = w.lml_k;
  a.nchunk = nchunk; a.chunk_len = chunk_len;
  a.bnd_m = w.bnd_m; a.bnd_P = w.bnd_P; a.unconverged = w.flag;
  a.from_bnd = 0; a.fixup = 0; a.warm = warm; a.delta = delta; a.patience = patience;
  const int64_t nfull = a.T / chunk_len;
  auto replay = [&](int64_t first, int64_t count) -> int {
    if (count <= 0) return PHYSS_OK;
    SeqFilterArgs r = a;
    r.chunk_first = first; r.chunk_count = count;
    return run_filter_any(st, d, m, disc_mode, nblk, hid, r);
  };
  // chunk 0 has no warm-up, chunks >= 1 all have exactly `warm` (<= chunk_len) steps of it: separate launches
  // keep the trip count uniform inside every warp
  rc = replay(0, nfull > 0 ? 1 : 0);
  if (rc) return rc;
  rc = replay(1, nfull - 1);
  if (rc) return rc;
  rc = replay(nfull, nchunk - nfull);
  if (rc) return rc;
  a.warm = 0; a.from_bnd = 1;
  for (int it = 0; it < polish && nchunk > 1; ++it) {
    const int64_t total = a.B * nchunk * (int64_t)(d * d + d);
    const int64_t want = (total + 255) / 256;
    ps_gather_bnd_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(a.mf, a.Pf, a.B, nchunk, chunk_len,
                                                                                     a.sbs, a.sts, d, w.bnd_m, w.bnd_P);
    rc = cuda_status(cudaGetLastError(), "ps_gather_bnd_kernel launch");
    if (rc) return rc;
    e = cudaMemsetAsync(w.flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return cuda_status(e, "pscan filter: flag reset");
    a.fixup = 1;
    rc = replay(1, nfull - 1);
    if (rc) return rc;
    if (nfull >= 1) { rc = replay(nfull, nchunk - nfull); if (rc) return rc; }
    a.fixup = 0;
  }
  {
    rc = ps_lml_sum(st, a.lml_k, a.B, a.T, a.sbs, a.sts, w.lml_partial, w.e0, a.B * nchunk * ps_filter_elem(d), a.lml);
    if (rc) return rc;
  }
  if (status_out) {
    e = cudaMemcpyAsync(status_out, w.flag, sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_status(e, "pscan filter: status copy");
  }
  return PHYSS_OK;
}

// Smoother counterpart: chunk c starts `warm` steps past its end from the filtered state there; full-state
// output only (the fix-up passes compare and carry the full smoothed state).
int pscan_smooth_spec(cudaStream_t st, int d, int mo, int disc_mode, int nblk, SeqSmoothArgs a, int64_t chunk_len,
                      int64_t warm, int polish, double delta, int patience, double* ws, int* status_out) {
  int rc = check_matern(d, disc_mode, nblk);
  if (rc) return rc;
  if (d > 32) return set_error(PHYSS_ERR_UNSUPPORTED, "pscan speculative mode: d <= 32 only");
  if (mo != 0) return set_error(PHYSS_ERR_UNSUPPORTED, "pscan speculative smoother: full_state output only");
  if (warm < 1 || warm > chunk_len) return set_error(PHYSS_ERR_BAD_ARG, "pscan speculative mode: need 1 <= warm <= chunk_len");
  if (polish < 1) polish = 1;
  const int64_t nchunk = ps_nchunk(a.T, chunk_len);
  PsWorkspace w = ps_carve(ws, a.B, a.T, d, chunk_len);
  cudaError_t e = cudaMemsetAsync(w.flag, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return cuda_status(e, "pscan smoother: flag reset");
  a.nchunk = nchunk; a.chunk_len = chunk_len;
  a.bnd_m = w.bnd_m; a.bnd_P = w.bnd_P; a.unconverged = w.flag;
  a.carry_last = 0; a.fixup = 0; a.warm = warm; a.delta = delta; a.patience = patience;
  auto replay = [&](int64_t first, int64_t count) -> int {
    if (count <= 0) return PHYSS_OK;
    SeqSmoothArgs r = a;
    r.chunk_first = first; r.chunk_count = count;
    return run_smooth_any(st, d, 0, disc_mode, nblk, r);
  };
  // uniform warm-up inside a launch: chunks [0, nchunk - 2) have >= chunk_len >= warm steps after them, the last
  // two chunks (shorter run-out / terminal condition) get their own launches
  const int64_t nmain = nchunk >= 2 ? nchunk - 2 : 0;
  rc = replay(0, nmain);
  if (rc) return rc;
  if (nchunk >= 2) { rc = replay(nchunk - 2, 1); if (rc) return rc; }
  rc = replay(nchunk - 1, 1);
  if (rc) return rc;
  a.warm = 0;
  for (int it = 0; it < polish && nchunk > 1; ++it) {
    const int64_t total = a.B * nchunk * (int64_t)(d * d + d);
    const int64_t want = (total + 255) / 256;
    ps_gather_bnd_next_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, st>>>(
        a.ms, a.Ps, a.B, nchunk, chunk_len, a.sbs, a.sts, d, w.bnd_m, w.bnd_P);
    rc = cuda_status(cudaGetLastError(), "ps_gather_bnd_next_kernel launch");
    if (rc) return rc;
    e = cudaMemsetAsync(w.flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return cuda_status(e, "pscan smoother: flag reset");
    a.fixup = 1;
    rc = replay(0, nchunk - 1);                              // all full-length chunks before the last one
    if (rc) return rc;
    a.fixup = 0;
  }
  if (status_out) {
    e = cudaMemcpyAsync(status_out, w.flag, sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_status(e, "pscan smoother: status copy");
  }
  return PHYSS_OK;
}

int pscan_filter_fold(cudaStream_t st, int d, int64_t B, int64_t K, const double* totals, const double* m0,
                      int64_t m0_bs, const double* P0, int64_t P0_bs, double* m_out, double* P_out) {
  const int G = ps_group_size(d);
  const PcLayout Lc = pc_layout(d);
  return PS_BY_G(run_filter_fold, st, totals, B, K, m0, m0_bs, P0, P0_bs, m_out, P_out, Lc);
}

// ------------------------------------------------------------------------------------ smoother entry
int pscan_smooth_local(cudaStream_t st, int d, int disc_mode, int nblk, SeqSmoothArgs a, int64_t chunk_len,
                       double* ws, double* total_out) {
  int rc = check_matern(d, disc_mode, nblk);
  if (rc) return rc;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  const int64_t nchunk = ps_nchunk(a.T, chunk_len);
  if (nchunk <= 1 && !total_out) return PHYSS_OK;
  PsWorkspace w = ps_carve(ws, a.B, a.T, d, chunk_len);
  a.nchunk = nchunk; a.chunk_len = chunk_len;
  const int G = ps_group_size(d);
  const PsLayout L = ps_layout(d, 1, given ? 0 : nblk, given, true);
  const SsLayout Ls = ss_layout(d);
  const int64_t nfull = a.T / chunk_len;
  const bool reg = d <= 4 && seq_supported(d, d, disc_mode, nblk);
  const bool rt = rt_supported(d, d) && !ps_force_grp();
  if (nfull > 0) {
    a.chunk_first = 0; a.chunk_count = nfull;
    rc = reg ? seq_smooth_summary(st, d, disc_mode, nblk, a, w.e0)
             : (rt ? rt_smooth_summary(st, d, disc_mode, nblk, a, w.e0) : PS_BY_G_GIVEN(run_smooth_summary, st, a, L, w.e0));
    if (rc) return rc;
  }
  if (nchunk > nfull) {
    a.chunk_first = nfull; a.chunk_count = nchunk - nfull;
    rc = reg ? seq_smooth_summary(st, d, disc_mode, nblk, a, w.e0)
             : (rt ? rt_smooth_summary(st, d, disc_mode, nblk, a, w.e0) : PS_BY_G_GIVEN(run_smooth_summary, st, a, L, w.e0));
    if (rc) return rc;
  }
  double* in = w.e0; double* out = w.e1;
  const bool fused = ps_fused_scan(d, nchunk) && a.B <= 0x7fffffff;
  if (fused) {
    for (int64_t stride = 1; stride < nchunk; stride *= 2) { double* t = in; in = out; out = t; }
    rc = run_scan_fused<true>(st, d, w.e0, in, a.B, nchunk, nchunk);
    if (rc) return rc;
  }
  for (int64_t stride = 1; !fused && stride < nchunk; stride *= 2) {
    rc = ps_reg_scan(d) ? run_smooth_scan_reg(st, d, in, out, a.B, nchunk, stride)
                        : PS_BY_G(run_smooth_scan, st, in, out, a.B, nchunk, stride, Ls);
    if (rc) return rc;
    double* t = in; in = out; out = t;
  }
  if (total_out) {
    const int64_t ns = ps_smooth_elem(d);
    rc = ps_copy_rows(st, total_out, ns, in, nchunk * ns, ns, a.B, "pscan smoother: copy of the range total");
    if (rc) return rc;
  }
  return PHYSS_OK;
}

// start_m / start_P: smoothed state one step past the end of this time range (carried from the next time
// shard); NULL = terminal condition, i.e. the range ends at the last step of the series.
int pscan_smooth_finish(cudaStream_t st, int d, int mo, int disc_mode, int nblk, SeqSmoothArgs a, int64_t chunk_len,
                        double* ws, const double* start_m, const double* start_P) {
  int rc = check_matern(d, disc_mode, nblk);
  if (rc) return rc;
  const int64_t nchunk = ps_nchunk(a.T, chunk_len);
  PsWorkspace w = ps_carve(ws, a.B, a.T, d, chunk_len);
  const int G = ps_group_size(d);
  const SsLayout Ls = ss_layout(d);
  cudaError_t e;
  if (!start_m) {
    // terminal: (mf, Pf)[T - 1]; its own step has dt = 0, so the RTS step reproduces smoothed = filtered
    const int64_t row = (a.T - 1) * a.sts;
    rc = ps_copy_rows(st, w.start_m, d, a.mf + row * d, a.sbs * d, d, a.B, "pscan smoother: copy of the terminal mean");
    if (rc) return rc;
    rc = ps_copy_rows(st, w.start_P, (int64_t)d * d, a.Pf + row * d * d, a.sbs * (int64_t)d * d, (int64_t)d * d, a.B,
                      "pscan smoother: copy of the terminal covariance");
    if (rc) return rc;
    start_m = w.start_m; start_P = w.start_P;
  }
  if (nchunk > 1) {
    const double* suffix = ps_scan_result(w, nchunk);
    rc = ps_reg_scan(d) ? run_smooth_apply_reg(st, d, suffix, a.B, nchunk, start_m, start_P, w.bnd_m, w.bnd_P)
                        : PS_BY_G(run_smooth_apply, st, suffix, a.B, nchunk, start_m, start_P, w.bnd_m, w.bnd_P, Ls);
    if (rc) return rc;
  } else {
    e = cudaMemcpyAsync(w.bnd_m, start_m, a.B * d * 8, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w.bnd_P, start_P, a.B * (size_t)d * d * 8, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_status(e, "pscan smoother: copy of the start state");
  }
  a.nchunk = nchunk; a.chunk_len = chunk_len;
  a.bnd_m = w.bnd_m; a.bnd_P = w.bnd_P; a.carry_last = 1; a.fixup = 0; a.unconverged = w.flag;
  const int64_t nfull = a.T / chunk_len;
  if (nfull > 0) {
    a.chunk_first = 0; a.chunk_count = nfull;
    rc = run_smooth_any(st, d, mo, disc_mode, nblk, a);
    if (rc) return rc;
  }
  if (nchunk > nfull) {
    a.chunk_first = nfull; a.chunk_count = nchunk - nfull;
    rc = run_smooth_any(st, d, mo, disc_mode, nblk, a);
    if (rc) return rc;
  }
  return PHYSS_OK;
}

int pscan_smooth_fold(cudaStream_t st, int d, int64_t B, int64_t K, const double* totals, const double* m0,
                      const double* P0, double* m_out, double* P_out) {
  const int G = ps_group_size(d);
  const SsLayout Ls = ss_layout(d);
  return PS_BY_G(run_smooth_fold, st, totals, B, K, m0, P0, m_out, P_out, Ls);
}

}  // namespace physs
