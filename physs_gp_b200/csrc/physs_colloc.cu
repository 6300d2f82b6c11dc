// physs_colloc.cu -- collocation (EKF) Kalman filter step, one thread per series: one latent with d <= 4, or systems of
// ODEs over 2 / 3 independent Matern-3/2 latents (d = 4 / 6: LotkaVolterra, LorenzSystem, transforms/pdes.py:818-1090).
//
// Reference: kf_predict_step(PDE, 'sequential'), computation/filters/kalman_filter.py:340-427, inside
// filter('sequential') (:439-485).  Per step:
//   1. LTI predict                     m_ = A m, P_ = A P A^T + Q                                  (:361-375)
//   2. residual and Jacobian at the PREDICTED mean: f = g(m_), H_jac = dg/dx(m_)                   (:378-379)
//   3. boundary update (optional)      H, R * 0, y = boundary_data[k], innovation H m_             (:382-391)
//   4. pseudo-observation update       H_jac, zero noise, y = pseudo observation, innovation f     (:395-414)
//   5. data update (observe_data)      H, R, y, innovation H m_                                    (:417-421)
//   the step's lml is that of the LAST update executed (the `ys` the step returns).
// The reference obtains H_jac with jax.jacfwd of PDE.forward_g (transforms/pdes.py:236-245) at a mean that lives
// INSIDE the recursion, so the residual has to be evaluated on chip.  It is described by a small table that covers
// the point-wise residuals the reference ships (Pendulum1D, DampedPendulum1D, SimpleODE, the u^3 - u reaction term
// of Allen-Cahn):   g_p(x, k) = w_p . x + sum_q coef_q phi_q(x[idx_q]) + forcing_p[k],   phi in {sin, cos, x^2, x^3}.
// The smoother of this model is the ordinary RTS smoother on the filtered moments (rts_smoother.py:108-150).
#include "physs_seq_impl.cuh"

namespace physs {

constexpr int kMaxTerms = 12;
constexpr int kMaxPc = 3, kMaxD = 8;

struct CollocArgs {
  int pc;                         // collocation outputs (PC template value)
  double w[kMaxPc][kMaxD];        // linear part  [pc][d]
  int n_terms;
  int t_out[kMaxTerms], t_kind[kMaxTerms], t_idx[kMaxTerms];
  double t_coef[kMaxTerms];
  const double* forcing;          // [pc, T] device, or NULL
  double y_pseudo[kMaxPc];        // 0, or NaN = this output is not collocated
  const double* boundary;         // step layout [.., M] device, or NULL
  int observe_data;
};

template <int D, int PC>
__device__ __forceinline__ void eval_residual(const CollocArgs& c, const double (&m)[D], int64_t k, int64_t T,
                                              double (&f)[PC], double (&Hj)[PC][D]) {
#pragma unroll
  for (int p = 0; p < PC; ++p) {
    double acc = (c.forcing && p < c.pc) ? c.forcing[p * T + k] : 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      acc = fma(c.w[p][j], m[j], acc);
      Hj[p][j] = c.w[p][j];
    }
    f[p] = acc;
  }
  for (int q = 0; q < c.n_terms; ++q) {                      // uniform over the grid
    const int idx = c.t_idx[q] & 255, idx2 = c.t_idx[q] >> 8, kind = c.t_kind[q], out = c.t_out[q];
    double x = 0.0, x2 = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {                            // no dynamic register indexing
      x = (idx == j) ? m[j] : x;
      x2 = (idx2 == j) ? m[j] : x2;
    }
    double val, der, der2 = 0.0;
    if (kind == PHYSS_RES_SIN) { val = sin(x); der = cos(x); }
    else if (kind == PHYSS_RES_COS) { val = cos(x); der = -sin(x); }
    else if (kind == PHYSS_RES_SQUARE) { val = x * x; der = 2.0 * x; }
    else if (kind == PHYSS_RES_PROD) { val = x * x2; der = x2; der2 = x; }
    else { val = x * x * x; der = 3.0 * x * x; }
    const double cf = c.t_coef[q];
#pragma unroll
    for (int p = 0; p < PC; ++p) {
      if (out == p) f[p] = fma(cf, val, f[p]);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        if (out == p && idx == j) Hj[p][j] = fma(cf, der, Hj[p][j]);
        if (kind == PHYSS_RES_PROD && out == p && idx2 == j) Hj[p][j] = fma(cf, der2, Hj[p][j]);
      }
    }
  }
}

template <int D, int S, int M, int PC, bool HID, bool GIVEN>
__global__ void __launch_bounds__(SeqBlock<D>::THREADS) seq_filter_colloc_kernel(const SeqFilterArgs p,
                                                                                 const CollocArgs c) {
  __shared__ __align__(16) double tiles[SeqBlock<D>::WARPS][RowTile<D * D>::SIZE];
  SeqWork wk;
  if (!seq_work<false>(p, wk)) return;
  constexpr int NB = D / S;
  const int64_t b = wk.b, T = wk.T;
  const bool active = wk.active;
  const int lane = threadIdx.x & 31;
  double* tile = tiles[threadIdx.x >> 5];
  const bool coal = (p.sbs == 1);
  const int64_t sts = p.sts;

  double m[D], P[D][D], Pinf[D][D], H[M][D], lam[NB];
  load_vec<D>(p.m0 + b * p.m0_bs, m);
  load_mat<D>(p.P0 + b * p.P0_bs, P);
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
  const int64_t row0 = b * p.sbs;
  const int64_t wrow0 = wk.b0 * p.sbs;
  const double* __restrict__ dtp = p.dt + b * p.dt_bs;
  const double* __restrict__ Yp = p.Y + row0 * M;
  const double* __restrict__ Bp = c.boundary ? c.boundary + row0 * M : nullptr;
  const double* __restrict__ Rp = p.R + b * p.R_bs;
  const double* __restrict__ Ap = GIVEN ? p.A + b * p.A_bs : nullptr;
  const double* __restrict__ Qp = GIVEN ? p.Q + b * p.Q_bs : nullptr;
  double* __restrict__ mfp = p.mf + row0 * D;
  double* __restrict__ Pfp = p.Pf + row0 * D * D;
  double* __restrict__ mfw = p.mf + wrow0 * D;
  double* __restrict__ Pfw = p.Pf + wrow0 * D * D;
  double* __restrict__ lkp = p.lml_k ? p.lml_k + row0 : nullptr;

  LmlAcc acc;
  for (int64_t k = 0; k < T; ++k) {
    double y[M], R[M][M];
    load_vec<M>(Yp + k * sts * M, y);
    load_mat<M>(Rp + k * p.R_ts, R);
    const double dt = dtp[k];
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
    // residual and its Jacobian at the predicted mean, BEFORE the boundary update (kalman_filter.py:378-379)
    double f[PC], Hj[PC][D];
    eval_residual<D, PC>(c, m, k, T, f, Hj);
    double det = 1.0, mahal = 0.0;
    int nobs = 0;
    if (Bp) {                                                  // uniform
      double yb[M], R0[M][M];
      load_vec<M>(Bp + k * sts * M, yb);
#pragma unroll
      for (int a = 0; a < M; ++a) {
#pragma unroll
        for (int cc = 0; cc < M; ++cc) R0[a][cc] = 0.0;
      }
      kf_update<D, M, HID>(m, P, H, R0, yb, p.jitter, det, mahal, nobs);
    }
    {
      double yp[PC], R0[PC][PC];
#pragma unroll
      for (int a = 0; a < PC; ++a) {
        yp[a] = c.y_pseudo[a];
#pragma unroll
        for (int cc = 0; cc < PC; ++cc) R0[a][cc] = 0.0;
      }
      kf_update<D, PC, false>(m, P, Hj, R0, yp, p.jitter, det, mahal, nobs, f);
    }
    if (c.observe_data) kf_update<D, M, HID>(m, P, H, R, y, p.jitter, det, mahal, nobs);
    acc.add(det, mahal, nobs);
    if (coal) {
      warp_store_rows<D>(mfw + k * sts * D, m, tile, lane, wk.nvalid);
      warp_store_rows<D * D>(Pfw + k * sts * D * D, flat<D>(P), tile, lane, wk.nvalid);
    } else if (active) {
      store_vec<D>(mfp + k * sts * D, m);
      store_mat<D>(Pfp + k * sts * D * D, P);
    }
    if (lkp && active) lkp[k * sts] = lml_term(det, mahal, nobs);
  }
  if (active) p.lml[b] = acc.value();
}

// NB = number of latents: systems observe one output per latent and carry 3 residuals for 3 latents
template <int D, int S, int NB, bool GIVEN>
static int colloc_launch(cudaStream_t st, const SeqFilterArgs& a, const CollocArgs& c, int m, bool hid) {
  const int64_t n = (a.B + 31) / 32 * 32;
  const int block = pick_block(n) < SeqBlock<D>::THREADS ? pick_block(n) : SeqBlock<D>::THREADS;
  const unsigned grid = (unsigned)((n + block - 1) / block);
  constexpr int PC = NB == 3 ? 3 : 2;
  if (hid && m == D) seq_filter_colloc_kernel<D, S, D, PC, true, GIVEN><<<grid, block, 0, st>>>(a, c);
  else if (m == NB) seq_filter_colloc_kernel<D, S, NB, PC, false, GIVEN><<<grid, block, 0, st>>>(a, c);
  else return set_error(PHYSS_ERR_UNSUPPORTED, "collocation filter: one observation per latent (or identity H with m == d)");
  return cuda_status(cudaGetLastError(), "seq_filter_colloc_kernel launch");
}

int colloc_filter(cudaStream_t st, int d, int m, int disc_mode, int nblk, bool hid, const SeqFilterArgs& a,
                  int pc, const double* res_w, int n_terms, const int32_t* t_out, const int32_t* t_kind,
                  const int32_t* t_idx, const double* t_coef, const double* forcing, const double* y_pseudo,
                  const double* boundary, int observe_data) {
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  // shapes: one latent (d = 2..4, given: one dense block), or -- Matern-3/2 latents only -- 2 / 3 latents of state dim 2
  int s = 0;
  if (given) s = (d >= 2 && d <= 4) || d == 6 ? d : 0;
  else if (d >= 2 && d <= 4 && nblk == 1) s = d;
  else if ((d == 4 || d == 6) && nblk == d / 2) s = 2;
  if (!s)
    return set_error(PHYSS_ERR_UNSUPPORTED, "collocation filter: one block with d = 2..4, or nblk = d / 2 blocks of size 2 with d = 4, 6");
  const int pcmax = d == 6 ? 3 : 2;
  if (pc < 1 || pc > pcmax || n_terms < 0 || n_terms > kMaxTerms || !res_w || !y_pseudo)
    return set_error(PHYSS_ERR_BAD_ARG, "collocation filter: 1..2 outputs (3 for three latents), at most 12 non-linear terms");
  CollocArgs c{};
  c.pc = pc;
  for (int p = 0; p < kMaxPc; ++p) {
    for (int j = 0; j < kMaxD; ++j) c.w[p][j] = (p < pc && j < d) ? res_w[p * d + j] : 0.0;
    // an absent output is a masked pseudo-observation: H row 0, y = NaN  ->  identity update
    c.y_pseudo[p] = (p < pc) ? y_pseudo[p] : nan("");
  }
  c.n_terms = n_terms;
  for (int q = 0; q < n_terms; ++q) {
    const int i1 = t_idx[q] & 255, i2 = t_idx[q] >> 8;
    const bool prod = t_kind[q] == PHYSS_RES_PROD;
    if (t_out[q] < 0 || t_out[q] >= pc || t_idx[q] < 0 || i1 >= d || (prod ? i2 >= d : i2 != 0) || t_kind[q] < 0 ||
        t_kind[q] > PHYSS_RES_PROD)
      return set_error(PHYSS_ERR_BAD_ARG, "collocation filter: bad residual term");
    c.t_out[q] = t_out[q]; c.t_kind[q] = t_kind[q]; c.t_idx[q] = t_idx[q]; c.t_coef[q] = t_coef[q];
  }
  c.forcing = forcing; c.boundary = boundary; c.observe_data = observe_data;
  if (given) {                // supplied transitions are dense: one d x d block whatever the number of latents
    if (d == 6) return colloc_launch<6, 6, 3, true>(st, a, c, m, hid);
    if (d == 4 && m == 2) return colloc_launch<4, 4, 2, true>(st, a, c, m, hid);
    if (d == 2) return colloc_launch<2, 2, 1, true>(st, a, c, m, hid);
    if (d == 3) return colloc_launch<3, 3, 1, true>(st, a, c, m, hid);
    return colloc_launch<4, 4, 1, true>(st, a, c, m, hid);
  }
  if (s == 2 && d == 4) return colloc_launch<4, 2, 2, false>(st, a, c, m, hid);
  if (s == 2 && d == 6) return colloc_launch<6, 2, 3, false>(st, a, c, m, hid);
  if (d == 2) return colloc_launch<2, 2, 1, false>(st, a, c, m, hid);
  if (d == 3) return colloc_launch<3, 3, 1, false>(st, a, c, m, hid);
  return colloc_launch<4, 4, 1, false>(st, a, c, m, hid);
}

}  // namespace physs
