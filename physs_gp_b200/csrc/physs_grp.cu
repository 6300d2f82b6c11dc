// physs_grp.cu -- general-dimension sequential Kalman filter / RTS smoother: one lane GROUP
// (8, 16 or 32 lanes) per series, every d x d block resident in shared memory, runtime (d, m).
//
// This is the path for state dims the register-resident kernels (physs_seq*.cu, d <= 4) do not cover:
// derivative-augmented / multi-latent states d = 5 .. ~40 with m = 1 .. d (BASELINE configs 3 and 5).
// Same reference semantics (kalman_filter.py:144-241,439-485; rts_smoother.py:48-106,162-192) and the
// same ABI entry points; physs_api.cu picks the implementation by size.
//
// Per step the streamed inputs (y_k, R_k, and for the smoother the filtered m_k, P_k; A_k/Q_k when the
// caller supplies them) are staged global -> shared with cp.async one step ahead, so HBM latency
// overlaps the algebra of the current step; outputs go shared -> global with consecutive lanes on
// consecutive addresses.
#include "physs_internal.h"
#include "physs_warp.cuh"

namespace physs {

using namespace grp;

struct GrpLayout {
  int d, m, mo, ld, ldm, nblk, s;
  // offsets in doubles inside one group's shared slab
  int P, A, Qm, W1, W2, W3, PfS[2], S, Sj, H, Ho, Rst[2], AQst[2][2];
  int vm, vmp, vv, vw, vrd, vy[2], vmf[2], vlam, vdm;
  int total;
};

static GrpLayout make_layout(int d, int m, int mo, int nblk, bool given, bool smoother) {
  GrpLayout L{};
  L.d = d; L.m = m; L.mo = mo; L.ld = d | 1; L.ldm = (m > 0 ? m : 1) | 1; L.nblk = nblk;
  L.s = (nblk > 0) ? d / nblk : d;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 1) & ~1; return o; };
  const int dd = d * L.ld;
  L.P = take(dd); L.A = take(dd); L.Qm = take(dd); L.W1 = take(dd); L.W2 = take(dd);
  if (smoother) {
    L.W3 = take(dd);
    L.PfS[0] = take(dd); L.PfS[1] = take(dd);
    L.vmf[0] = take(d); L.vmf[1] = take(d);
    L.Ho = take((mo > 0 ? mo : 0) * L.ld);
  } else {
    L.S = take(m * L.ldm); L.Sj = take(m * L.ldm);
    L.H = take(m * L.ld);
    L.Rst[0] = take(m * L.ldm); L.Rst[1] = take(m * L.ldm);
    L.vy[0] = take(m); L.vy[1] = take(m);
    L.vv = take(m); L.vw = take(m);
  }
  if (given) {
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L.AQst[a][b] = take(dd);
  }
  L.vm = take(d); L.vmp = take(d); L.vdm = take(d); L.vrd = take(d > m ? d : m);
  L.vlam = take(nblk > 0 ? nblk : 1);
  L.total = off;
  return L;
}

// ------------------------------------------------------------------------------------------ filter
// relative agreement of the shared-memory (m, P) with the stored global one (chunk-mode fix-up);
// evaluated redundantly by every lane of the group
__device__ __forceinline__ bool grp_agrees(const double* mv, const double* P, int ld, int d,
                                           const double* om, const double* oP, double delta) {
  double dP = 0.0, sP = 0.0, dm = 0.0, sm = 0.0;
  for (int i = 0; i < d; ++i) {
    const double omi = om[i];
    dm = fmax(dm, fabs(mv[i] - omi));
    sm = fmax(sm, fabs(omi));
    for (int j = 0; j < d; ++j) {
      const double o = oP[i * d + j];
      dP = fmax(dP, fabs(P[i * ld + j] - o));
      sP = fmax(sP, fabs(o));
    }
  }
  return (dP <= delta * sP) && (dm <= delta * sm || dm * dm <= delta * delta * sP);
}

// (series, chunk) owned by a group.  Plain mode: chunk 0 = the whole series.
struct GrpWork {
  int64_t b, c, v, t0, T, Tfull;
  bool active, chunked;
};

template <typename Args>
__device__ __forceinline__ GrpWork grp_work(const Args& p, int64_t gid) {
  GrpWork w;
  w.chunked = p.nchunk > 0;
  const int64_t per = w.chunked ? p.chunk_count : 1;
  const int64_t n = p.B * per;
  w.active = gid < n;
  const int64_t g = w.active ? gid : n - 1;     // inactive groups shadow the last one (no stores)
  w.b = g / per;
  w.c = w.chunked ? p.chunk_first + g % per : 0;
  w.v = w.chunked ? w.b * p.nchunk + w.c : w.b;
  w.t0 = w.chunked ? w.c * p.chunk_len : 0;
  w.Tfull = p.T;
  w.T = w.chunked ? ((p.chunk_len < w.Tfull - w.t0) ? p.chunk_len : (w.Tfull - w.t0)) : w.Tfull;
  return w;
}

template <int G, bool GIVEN>
__global__ void grp_filter_kernel(const SeqFilterArgs p, const GrpLayout L, const bool hid) {
  extern __shared__ __align__(16) double smem[];
  const int groups_per_block = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const GrpWork wk = grp_work(p, (int64_t)blockIdx.x * groups_per_block + g_in_block);
  const bool active = wk.active, chunked = wk.chunked;
  const int64_t bb = wk.b, b = wk.b, vs = wk.v, t0 = wk.t0, Tfull = wk.Tfull;
  const int gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = L.d, m = L.m, ld = L.ld, ldm = L.ldm, s = L.s;
  const int64_t T = wk.T;

  double* P = sm + L.P; double* A = sm + L.A; double* Qm = sm + L.Qm;
  double* W1 = sm + L.W1; double* W2 = sm + L.W2;
  double* S = sm + L.S; double* Sj = sm + L.Sj; double* H = sm + L.H;
  double* mv_ = sm + L.vm; double* mp = sm + L.vmp; double* v = sm + L.vv; double* w = sm + L.vw;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;

  if (chunked && p.from_bnd) {
    g2s<G>(P, ld, p.bnd_P + vs * d * d, d, d);
    for (int i = gl; i < d; i += G) mv_[i] = p.bnd_m[vs * d + i];
  } else {
    g2s<G>(P, ld, p.P0 + bb * p.P0_bs, d, d);
    for (int i = gl; i < d; i += G) mv_[i] = p.m0[bb * p.m0_bs + i];
  }
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
  double* mfp = p.mf + row0 * d;
  double* Pfp = p.Pf + row0 * d * d;
  double* lkp = p.lml_k ? p.lml_k + row0 : nullptr;
  int streak = 0;
  bool done = false;

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

  LmlAcc acc;
  stage(0);
  double dt_n = dtp[0];
  for (int64_t k = 0; k < T; ++k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    cp_async_wait_all();
    __syncwarp();
    if (k + 1 < T) { stage(k + 1); dt_n = dtp[k + 1]; }
    const double* y = sm + L.vy[st];
    const double* R = sm + L.Rst[st];
    const double* Ak = A;
    const double* Qk = Qm;
    int bsA = s;
    // ---- predict
    if (GIVEN) {
      Ak = sm + L.AQst[st][0];
      Qk = sm + L.AQst[st][1];
      bsA = 0;
      mv<G, false>(mp, Ak, ld, mv_, d, d, nullptr, 1.0);
      mm<G, false, false>(W2, ld, Ak, ld, P, ld, d, d, d, nullptr, 0, 1.0);        // A P
      __syncwarp();
      mm<G, false, true>(P, ld, W2, ld, Ak, ld, d, d, d, Qk, ld, 1.0);               // A P A^T + Q
    } else {
      matern_A<G>(A, ld, d, s, L.nblk, lam, dt);
      for (int idx = gl; idx < d * d; idx += G) {                                   // dP = P - Pinf
        const int i = idx / d, j = idx - i * d;
        W1[i * ld + j] = P[i * ld + j] - Qm[i * ld + j];
      }
      __syncwarp();
      mv<G, false>(mp, A, ld, mv_, d, d, nullptr, 1.0, bsA);
      mm<G, false, false>(W2, ld, A, ld, W1, ld, d, d, d, nullptr, 0, 1.0, bsA, 0);   // A dP
      __syncwarp();
      mm<G, false, true>(P, ld, W2, ld, A, ld, d, d, d, Qm, ld, 1.0, 0, bsA);         // Pinf + A dP A^T
    }
    __syncwarp();
    // ---- update: HP = M H P_ (rows of missing obs zeroed) -> W1 [m x d]
    for (int idx = gl; idx < m * d; idx += G) {
      const int a = idx / d, j = idx - a * d;
      double accv;
      if (hid) {
        accv = P[a * ld + j];
      } else {
        accv = 0.0;
        for (int l = 0; l < d; ++l) accv = fma(H[a * ld + l], P[l * ld + j], accv);
      }
      const double ya = y[a];
      W1[a * ld + j] = (ya != ya) ? 0.0 : accv;
    }
    for (int a = gl; a < m; a += G) {
      double mu;
      if (hid) {
        mu = mp[a];
      } else {
        mu = 0.0;
        for (int l = 0; l < d; ++l) mu = fma(H[a * ld + l], mp[l], mu);
      }
      const double ya = y[a];
      v[a] = (ya != ya) ? 0.0 : (ya - mu);
    }
    __syncwarp();
    // S = M H P_ H^T M + R ; Sj = S + jitter I
    for (int idx = gl; idx < m * m; idx += G) {
      const int a = idx / m, c = idx - a * m;
      double accv;
      if (hid) {
        accv = W1[a * ld + c];
      } else {
        accv = 0.0;
        for (int l = 0; l < d; ++l) accv = fma(W1[a * ld + l], H[c * ld + l], accv);
      }
      const double yc = y[c];
      accv = (yc != yc) ? 0.0 : accv;
      const double sv = accv + R[a * ldm + c];
      S[a * ldm + c] = sv;
      Sj[a * ldm + c] = sv + (a == c ? p.jitter : 0.0);
    }
    __syncwarp();
    chol<G>(Sj, ldm, m, rd);
    chol_solve<G>(Sj, ldm, m, rd, W1, ld, d);                 // W1 <- K^T [m x d]
    __syncwarp();
    // m = m_ + K v
    for (int i = gl; i < d; i += G) {
      double accv = mp[i];
      for (int a = 0; a < m; ++a) accv = fma(W1[a * ld + i], v[a], accv);
      mv_[i] = accv;
    }
    // KS = K S  -> W2 [d x m] (ld ldm)
    for (int idx = gl; idx < d * m; idx += G) {
      const int i = idx / m, c = idx - i * m;
      double accv = 0.0;
      for (int a = 0; a < m; ++a) accv = fma(W1[a * ld + i], S[a * ldm + c], accv);
      W2[i * ldm + c] = accv;
    }
    __syncwarp();
    // P -= KS K^T
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      double accv = P[i * ld + j];
      for (int c = 0; c < m; ++c) accv = fma(-W2[i * ldm + c], W1[c * ld + j], accv);
      P[i * ld + j] = accv;
    }
    // lml: un-jittered S with missing rows/cols -> identity
    int nobs = 0;
    for (int a = 0; a < m; ++a) nobs += (y[a] != y[a]) ? 0 : 1;
    for (int idx = gl; idx < m * m; idx += G) {
      const int a = idx / m, c = idx - a * m;
      const bool keep = !(y[a] != y[a]) && !(y[c] != y[c]);
      Sj[a * ldm + c] = keep ? S[a * ldm + c] : (a == c ? 1.0 : 0.0);
    }
    for (int a = gl; a < m; a += G) w[a] = v[a];
    __syncwarp();
    const double det = chol<G>(Sj, ldm, m, rd);
    chol_solve<G>(Sj, ldm, m, rd, w, 1, 1);
    __syncwarp();
    double mahal = 0.0;
    for (int a = 0; a < m; ++a) mahal = fma(v[a], w[a], mahal);
    if (!chunked) acc.add(det, mahal, nobs);
    // ---- outputs (fix-up: compare with what is stored before overwriting it)
    if (chunked && p.fixup && active && !done)
      streak = grp_agrees(mv_, P, ld, d, mfp + k * sts * d, Pfp + k * sts * d * d, p.delta) ? streak + 1 : 0;
    __syncwarp();
    if (active && !done) {
      for (int i = gl; i < d; i += G) mfp[k * sts * d + i] = mv_[i];
      s2g<G>(Pfp + k * sts * d * d, P, ld, d, d);
      if (lkp && gl == 0) lkp[k * sts] = lml_term(det, mahal, nobs);
    }
    if (chunked && p.fixup) {
      if (streak >= p.patience) done = true;
      if (__all_sync(0xffffffffu, done || !active)) break;   // every group of the warp converged
    }
    __syncwarp();
  }
  if (chunked) {
    if (p.fixup && active && !done && gl == 0) atomicOr(p.unconverged, 1);
  } else if (active && gl == 0) {
    p.lml[b] = acc.value();
  }
}

// ---------------------------------------------------------------------------------------- smoother
template <int G, bool GIVEN>
__global__ void grp_smooth_kernel(const SeqSmoothArgs p, const GrpLayout L) {
  extern __shared__ __align__(16) double smem[];
  const int groups_per_block = blockDim.x / G;
  const int g_in_block = threadIdx.x / G;
  const GrpWork wk = grp_work(p, (int64_t)blockIdx.x * groups_per_block + g_in_block);
  const bool active = wk.active, chunked = wk.chunked;
  const int64_t bb = wk.b, vs = wk.v, t0 = wk.t0, Tfull = wk.Tfull;
  const int gl = Lanes<G>::gl();
  double* sm = smem + (size_t)g_in_block * L.total;
  const int d = L.d, mo = L.mo, ld = L.ld, s = L.s;
  const int mp_ = (mo == 0) ? d : mo;
  const int64_t T = wk.T;

  double* Ps = sm + L.P; double* A = sm + L.A; double* Qm = sm + L.Qm;
  double* W1 = sm + L.W1; double* W2 = sm + L.W2; double* W3 = sm + L.W3; double* Ho = sm + L.Ho;
  double* ms = sm + L.vm; double* mpred = sm + L.vmp; double* dm = sm + L.vdm;
  double* rd = sm + L.vrd; double* lam = sm + L.vlam;

  if (!GIVEN) {
    g2s<G>(Qm, ld, p.Pinf + bb * p.Pinf_bs, d, d);
    for (int i = gl; i < L.nblk; i += G) lam[i] = p.lam[bb * p.lam_bs + i];
  }
  if (mo > 0) g2s<G>(Ho, ld, p.Hout, mo, d);
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
        s2g<G>(Psp + k * sts * d * d, Ps, ld, d, d);
      }
    } else {
      // W1 <- Hout Ps  [mo x d];  out = W1 Hout^T
      mm<G, false, false>(W1, ld, Ho, ld, Ps, ld, mo, d, d, nullptr, 0, 1.0);
      __syncwarp();
      if (active) {
        mm<G, false, true>(Psp + k * sts * mo * mo, mo, W1, ld, Ho, ld, mo, d, mo, nullptr, 0, 1.0);
        mv<G, false>(msp + k * sts * mo, Ho, ld, ms, mo, d, nullptr, 1.0);
      }
    }
    __syncwarp();
  };

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

  // plain mode: the last step is terminal.  Chunk mode: every step is an RTS step from the carried
  // state of the next chunk's first step; the last chunk carries its own last filtered state over
  // dt = 0, which reproduces the terminal condition (see physs_seq_impl.cuh).
  const bool carried = chunked && (wk.c < p.nchunk - 1 || p.carry_last);
  if (carried) {
    g2s<G>(Ps, ld, p.bnd_P + vs * d * d, d, d);
    for (int i = gl; i < d; i += G) ms[i] = p.bnd_m[vs * d + i];
  } else {
    g2s<G>(Ps, ld, Pfp + (T - 1) * sts * d * d, d, d);
    for (int i = gl; i < d; i += G) ms[i] = mfp[(T - 1) * sts * d + i];
  }
  __syncwarp();
  int64_t kstart = T - 1;
  if (!chunked) {
    emit(T - 1);
    kstart = T - 2;
  }
  double dt_n = 0.0;
  if (kstart >= 0) { stage(kstart); dt_n = dtp[kstart]; }
  for (int64_t k = kstart; k >= 0; --k) {
    const int st = (int)(k & 1);
    const double dt = dt_n;
    cp_async_wait_all();
    __syncwarp();
    if (k >= 1) { stage(k - 1); dt_n = dtp[k - 1]; }
    const double* mf = sm + L.vmf[st];
    const double* Pf = sm + L.PfS[st];
    const double* Ak = A;
    int bsA = s;
    if (GIVEN) {
      Ak = sm + L.AQst[st][0];
      const double* Qk = sm + L.AQst[st][1];
      bsA = 0;
      mv<G, false>(mpred, Ak, ld, mf, d, d, nullptr, 1.0);
      mm<G, false, false>(W1, ld, Ak, ld, Pf, ld, d, d, d, nullptr, 0, 1.0);          // C = A Pf
      __syncwarp();
      mm<G, false, true>(W2, ld, W1, ld, Ak, ld, d, d, d, Qk, ld, 1.0);                 // Pp = C A^T + Q
    } else {
      matern_A<G>(A, ld, d, s, L.nblk, lam, dt);
      __syncwarp();
      mv<G, false>(mpred, A, ld, mf, d, d, nullptr, 1.0, bsA);
      mm<G, false, false>(W1, ld, A, ld, Pf, ld, d, d, d, nullptr, 0, 1.0, bsA, 0);     // C = A Pf
      mm<G, false, false>(W3, ld, A, ld, Qm, ld, d, d, d, nullptr, 0, 1.0, bsA, 0);     // A Pinf
      __syncwarp();
      for (int idx = gl; idx < d * d; idx += G) {
        const int i = idx / d, j = idx - i * d;
        W3[i * ld + j] = W1[i * ld + j] - W3[i * ld + j];
      }
      __syncwarp();
      mm<G, false, true>(W2, ld, W3, ld, A, ld, d, d, d, Qm, ld, 1.0, 0, bsA);          // Pp
    }
    __syncwarp();
    // dP = Ps - Pp -> W3 ; dm = ms - mpred ; Pp += jitter I (in place, then factor)
    for (int idx = gl; idx < d * d; idx += G) {
      const int i = idx / d, j = idx - i * d;
      const double pp = W2[i * ld + j];
      W3[i * ld + j] = Ps[i * ld + j] - pp;
      if (i == j) W2[i * ld + j] = pp + p.jitter;
    }
    for (int i = gl; i < d; i += G) dm[i] = ms[i] - mpred[i];
    __syncwarp();
    chol<G>(W2, ld, d, rd);
    chol_solve<G>(W2, ld, d, rd, W1, ld, d);                  // W1 <- X = (Pp + jit)^-1 A Pf = G^T
    __syncwarp();
    // ms = mf + G dm = mf + X^T dm
    mv<G, true>(ms, W1, ld, dm, d, d, mf, 1.0);
    // W2 <- G dP = X^T dP
    mm<G, true, false>(W2, ld, W1, ld, W3, ld, d, d, d, nullptr, 0, 1.0);
    __syncwarp();
    // Ps = Pf + (G dP) G^T = Pf + W2 X
    mm<G, false, false>(Ps, ld, W2, ld, W1, ld, d, d, d, Pf, ld, 1.0);
    __syncwarp();
    if (chunked && p.fixup && mo == 0 && active && !done)
      streak = grp_agrees(ms, Ps, ld, d, msp + k * sts * d, Psp + k * sts * d * d, p.delta) ? streak + 1 : 0;
    __syncwarp();
    emit(k);
    if (chunked && p.fixup) {
      if (streak >= p.patience) done = true;
      if (__all_sync(0xffffffffu, done || !active)) break;
    }
  }
  if (chunked && p.fixup && active && !done && gl == 0) atomicOr(p.unconverged, 1);
}

// ---------------------------------------------------------------------------------------- dispatch
static int group_size(int d) { return d <= 8 ? 8 : (d <= 16 ? 16 : 32); }

bool grp_supported(int d, int m) {
  if (d < 1 || m < 1 || m > d) return false;
  GrpLayout L = make_layout(d, m, d, 1, true, true);
  return (size_t)L.total * sizeof(double) <= 200 * 1024;
}

template <int G, bool GIVEN>
static int run_filter(cudaStream_t st, const SeqFilterArgs& a, const GrpLayout& L, bool hid) {
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 200 * 1024) threads /= 2;
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024)
    return set_error(PHYSS_ERR_UNSUPPORTED, "state dimension too large for the shared-memory path");
  const int gpb = threads / G;
  const int64_t ngroups = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  const int64_t grid = (ngroups + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(grp_filter_kernel<G, GIVEN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(grp_filter_kernel)");
  grp_filter_kernel<G, GIVEN><<<(unsigned)grid, threads, smem, st>>>(a, L, hid);
  return cuda_status(cudaGetLastError(), "grp_filter_kernel launch");
}

template <int G, bool GIVEN>
static int run_smooth(cudaStream_t st, const SeqSmoothArgs& a, const GrpLayout& L) {
  const size_t per_group = (size_t)L.total * sizeof(double);
  int threads = 128;
  while (threads > 32 && per_group * (threads / G) > 200 * 1024) threads /= 2;
  const size_t smem = per_group * (threads / G);
  if (smem > 200 * 1024)
    return set_error(PHYSS_ERR_UNSUPPORTED, "state dimension too large for the shared-memory path");
  const int gpb = threads / G;
  const int64_t ngroups = a.B * (a.nchunk > 0 ? a.chunk_count : 1);
  const int64_t grid = (ngroups + gpb - 1) / gpb;
  cudaError_t e = cudaFuncSetAttribute(grp_smooth_kernel<G, GIVEN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(grp_smooth_kernel)");
  grp_smooth_kernel<G, GIVEN><<<(unsigned)grid, threads, smem, st>>>(a, L);
  return cuda_status(cudaGetLastError(), "grp_smooth_kernel launch");
}

int grp_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool h_identity,
               const SeqFilterArgs& a) {
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  if (!given) {
    const int s = d / nblk;
    if (s < 1 || s > 4 || s * nblk != d)
      return set_error(PHYSS_ERR_UNSUPPORTED, "DISC_MATERN needs equal blocks of size 1..4");
  }
  GrpLayout L = make_layout(d, m, 0, given ? 0 : nblk, given, false);
  const int G = group_size(d);
#define RUN(G_) (given ? run_filter<G_, true>(st, a, L, h_identity) : run_filter<G_, false>(st, a, L, h_identity))
  if (G == 8) return RUN(8);
  if (G == 16) return RUN(16);
  return RUN(32);
#undef RUN
}

int grp_smooth(cudaStream_t st, int d, int mo, int disc_mode, int nblk, const SeqSmoothArgs& a) {
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  if (!given) {
    const int s = d / nblk;
    if (s < 1 || s > 4 || s * nblk != d)
      return set_error(PHYSS_ERR_UNSUPPORTED, "DISC_MATERN needs equal blocks of size 1..4");
  }
  GrpLayout L = make_layout(d, 1, mo, given ? 0 : nblk, given, true);
  const int G = group_size(d);
#define RUN(G_) (given ? run_smooth<G_, true>(st, a, L) : run_smooth<G_, false>(st, a, L))
  if (G == 8) return RUN(8);
  if (G == 16) return RUN(16);
  return RUN(32);
#undef RUN
}

}  // namespace physs
