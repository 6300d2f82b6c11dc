/* physs_b200.h -- C ABI of libphyss_b200.so: the B200 (sm_100a) state-space inference hot path.
 *
 * Drop-in boundary.  The reference (jonathanfrennert/physs_gp, pure Python/JAX) has no FFI; its
 * operator API for this path is the string-keyed registry in src/lib/stgp/dispatch.py:133-189:
 *     evoke('filter',   filter_type)  -> filter(data, prior, lik_mat, Y, X_t, X_s, dt, lik_cov_flag,
 *                                               train_test_mask, train_index)
 *                                        (computation/filters/kalman_filter.py:439-485, called at :541)
 *     evoke('smoother', filter_type)  -> smoother(data, model, filter_res, dt, X_t, X_s, full_state)
 *                                        (computation/filters/rts_smoother.py:162-192, called at :215)
 * A backend registered as filter_type='b200' evaluates the (tiny, T-independent) prior quantities on
 * the host and makes ONE call into this library per filter / smoother pass.  Each entry point below
 * cites the reference function it replaces.  INTEGRATION.md shows the jax.ffi / ctypes stubs.
 *
 * Conventions
 *   - All pointers are DEVICE pointers to fp64 (C row-major, reference element order) unless noted;
 *     `stream` is a cudaStream_t passed as void*.  Calls are stream-ordered and never synchronise.
 *   - Matrix / vector arrays (everything except dt, lam, lml) must be 16-byte aligned (any CUDA or
 *     torch allocation is); the kernels use 128-bit accesses.
 *   - The caller owns every buffer; the library allocates nothing user-visible.
 *   - A leading batch axis B (independent series / spatial blocks / latent functions) is added in
 *     front of every reference array.  A `*_bstride` argument is the element stride between series
 *     for that array; 0 means "shared by all series".
 *   - Per-step arrays (Y, mf, Pf, ms, Ps, lml_k) share one pair of "step strides": the row of series b
 *     at step k of an array with n doubles per step starts at base + (b*step_bstride + k*step_tstride)*n.
 *     (0, 0) or (T, 1) = batch-major [B][T][n] (what jax.vmap over axis 0 of the reference returns);
 *     (1, B) = time-major [T][B][n] (vmap with out_axes=1) -- the layout the kernels are fastest in: the
 *     32 series of a warp then read / write one contiguous span per step with coalesced 16-byte accesses.
 *   - NaN in Y marks a missing observation (utils/nan_utils.py:13-20).
 *   - Numerical failure (non-PD Cholesky) writes NaN and still returns 0, as jnp.linalg.cholesky does;
 *     the reference's NaN guard (trainers/natgrad_trainer.py:257-285) keeps working.
 *   - Return value: 0 = ok; PHYSS_ERR_* otherwise (bad arguments / unsupported size / CUDA launch
 *     failure).  No exceptions cross the ABI.
 *   - jitter values are runtime arguments (settings.py:63-64 are read at trace time in the reference).
 */
#ifndef PHYSS_B200_H_
#define PHYSS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHYSS_OK 0
#define PHYSS_ERR_BAD_ARG 1
#define PHYSS_ERR_UNSUPPORTED 2
#define PHYSS_ERR_CUDA 3

/* Discretisation modes (how A_k = expm(F dt_k) and Q_k reach the kernel). */
#define PHYSS_DISC_GIVEN 0  /* caller supplies A[.,T,d,d] and Q[.,T,d,d]                                  */
#define PHYSS_DISC_MATERN 1 /* block-diagonal stack of nblk Matern-(s-1/2) blocks of equal size s=d/nblk, */
                            /* closed-form expm (kernels/ss_utils.py:6-10, kernels/matern.py:152-177,    */
                            /* 306-329) from lam[., nblk] = sqrt(2 nu)/lengthscale, Q_k = Pinf - A Pinf A^T */
                            /* (kernels/kernel.py:207-209) with block-diagonal Pinf[., d, d].             */
#define PHYSS_DISC_IWP 2    /* one integrated-Wiener block IWP(q), q = d - 1 in 1..3 (WienerVelocity,     */
                            /* kernels/wiener.py:60-149): A_k[i][j] = dt^(j-i)/(j-i)!, Q_k[i][j] =        */
                            /* lam dt^(2q+1-i-j) / ((2q+1-i-j)(q-i)!(q-j)!) on chip; lam[., 1] = spectral  */
                            /* density (the kernel's variance); Pinf unused (may be NULL).  Sequential     */
                            /* filter / smoother only; larger or stacked IWP priors go through DISC_GIVEN. */

/* ABI version (bumped on any signature change). */
int physs_abi_version(void);

/* Human-readable description of the last non-zero status on this thread. */
const char* physs_last_error(void);

/* 1 if a kernel specialisation exists for this (d, m, disc_mode, nblk), else 0. */
int physs_kf_supported(int32_t d, int32_t m, int32_t disc_mode, int32_t nblk);

/* Number of independent series ONE full wave of the smoother kernel for this shape keeps resident on the current
 * device (occupancy x series per block x SM count), or 0 when the shape has no fixed-wave kernel.  Every series is
 * walked by one thread / lane group for the whole launch, so a batch that is a whole number of waves never ends
 * on a half-empty GPU: callers that split a large batch (the reference has no batch axis; jax.vmap callers pick
 * the chunking) should use multiples of this.  Host-only query, no launch, no stream. */
int64_t physs_kf_wave_series(int32_t d, int32_t disc_mode, int32_t nblk);

/* Sequential Kalman filter over B independent series.
 * Replaces filter('sequential') + kf_predict_step(LTI_SDE) + kf_update_step
 * (computation/filters/kalman_filter.py:439-485, 214-241, 144-211).
 *   dt   [., T]      dt[0] = 0, dt[k] = t_k - t_{k-1}  (filter_loop, kalman_filter.py:515)
 *   m0   [., d]      initial mean   (prior.m_inf)
 *   P0   [., d, d]   initial covariance (prior.P_inf)
 *   Pinf [., d, d]   stationary covariance (DISC_MATERN only; may alias P0)
 *   H    [., m, d]   measurement matrix (prior.H)
 *   Y    [B, T, m]   observations, NaN = missing
 *   R    [., ., m, m] observation covariance; R_bstride / R_tstride in elements (0 = broadcast)
 * Outputs
 *   mf   [B, T, d], Pf [B, T, d, d]   filtered moments (reference: {'m': [T,d,1], 'P': [T,d,d]})
 *   lml  [B]                          sum_k lml_k
 *   lml_k [B, T] or NULL              per-step terms
 */
int physs_kf_filter_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                        int32_t d, int32_t m,
                        int32_t disc_mode, int32_t nblk,
                        const double* A, int64_t A_bstride,
                        const double* Q, int64_t Q_bstride,
                        const double* lam, int64_t lam_bstride,
                        const double* dt, int64_t dt_bstride,
                        const double* Pinf, int64_t Pinf_bstride,
                        const double* m0, int64_t m0_bstride,
                        const double* P0, int64_t P0_bstride,
                        const double* H, int64_t H_bstride,
                        const double* Y,
                        const double* R, int64_t R_bstride, int64_t R_tstride,
                        double jitter,
                        double* mf, double* Pf, double* lml, double* lml_k);

/* Sequential RTS smoother over B independent series.
 * Replaces smoother('sequential') + rts_step_wrapper(LTI_SDE) + rts_smoother_step
 * (computation/filters/rts_smoother.py:162-192, 69-106, 48-65).
 *   dt   [., T]      dt[k] = t_{k+1} - t_k, dt[T-1] = 0  (smoother_loop, rts_smoother.py:209-211)
 *   mf, Pf           filtered moments from physs_kf_filter_f64
 *   Hout [mo, d] or NULL; mo = 0 / NULL means full_state=True (H = I, rts_smoother.py:28-31)
 * Outputs
 *   ms [B, T, mo'], Ps [B, T, mo', mo']   with mo' = (mo == 0 ? d : mo)
 */
int physs_rts_smooth_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                         int32_t d,
                         int32_t disc_mode, int32_t nblk,
                         const double* A, int64_t A_bstride,
                         const double* Q, int64_t Q_bstride,
                         const double* lam, int64_t lam_bstride,
                         const double* dt, int64_t dt_bstride,
                         const double* Pinf, int64_t Pinf_bstride,
                         const double* mf, const double* Pf,
                         const double* Hout, int32_t mo,
                         double jitter,
                         double* ms, double* Ps);

/* Filter + smoother in one call (`physs_kf_filter_smooth_f64`): physs_kf_filter_f64
 * followed by physs_rts_smooth_f64 on the same stream with the filter's outputs -- what
 * `BASE_SDE_GP.filter_and_smooth` (models/sde_gp.py:212-302) does with two dispatched calls; one FFI custom call
 * instead of two.  dt_smooth [., T] is the smoother's convention (dt[k] = t_{k+1} - t_k, dt[T-1] = 0). */
int physs_kf_filter_smooth_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                               int32_t d, int32_t m,
                               int32_t disc_mode, int32_t nblk,
                               const double* A, int64_t A_bstride,
                               const double* Q, int64_t Q_bstride,
                               const double* lam, int64_t lam_bstride,
                               const double* dt, int64_t dt_bstride,
                               const double* Pinf, int64_t Pinf_bstride,
                               const double* m0, int64_t m0_bstride,
                               const double* P0, int64_t P0_bstride,
                               const double* H, int64_t H_bstride,
                               const double* Y,
                               const double* R, int64_t R_bstride, int64_t R_tstride,
                               double jitter,
                               const double* A_smooth, const double* Q_smooth,
                               const double* dt_smooth, int64_t dt_smooth_bstride,
                               const double* Hout, int32_t mo,
                               double* mf, double* Pf, double* lml, double* lml_k, double* ms, double* Ps);

/* Filter + smoother in one call WITHOUT the filtered moments as an output: what `BASE_SDE_GP.filter_and_smooth`
 * (models/sde_gp.py:212-302) needs when the caller only wants the smoothed (projected) posterior and the lml -- the
 * reference's filter_loop (kalman_filter.py:487-547) still materialises every filtered (m, P) for smoother_loop
 * (rts_smoother.py:194-219) to read back.  Here the hand-over stays in a caller-supplied workspace as packed rows
 * [m (d) | upper triangle of P (d (d + 1) / 2)] (the update leaves P bitwise symmetric, so the smoother sees exactly
 * the (m, P) the two-call path stores): 14 instead of 20 doubles per step each way at d = 4, i.e. 264 instead of
 * 360 B per state-step with a projected scalar output.  Same arithmetic, operation for operation: lml and the
 * full-state (ms, Ps) are BITWISE those of physs_kf_filter_smooth_f64; projected outputs agree to 1e-14 (two
 * instantiations of one source, nvcc picks one contraction differently).
 * Register kernels only: d <= 4 (physs_kf_filter_smooth_packed_supported), time-major steps
 * (step_bstride == 1, step_tstride >= B); anything else returns PHYSS_ERR_UNSUPPORTED and the caller takes the
 * two-output call.  ws: device memory, 16-byte aligned, >= physs_kf_filter_smooth_packed_ws_bytes(...) bytes. */
int physs_kf_filter_smooth_packed_supported(int32_t d, int32_t m, int32_t disc_mode, int32_t nblk);
int64_t physs_kf_filter_smooth_packed_ws_bytes(int64_t B, int64_t T, int64_t step_tstride, int32_t d);
int physs_kf_filter_smooth_packed_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                                      int32_t d, int32_t m,
                                      int32_t disc_mode, int32_t nblk,
                                      const double* A, int64_t A_bstride,
                                      const double* Q, int64_t Q_bstride,
                                      const double* lam, int64_t lam_bstride,
                                      const double* dt, int64_t dt_bstride,
                                      const double* Pinf, int64_t Pinf_bstride,
                                      const double* m0, int64_t m0_bstride,
                                      const double* P0, int64_t P0_bstride,
                                      const double* H, int64_t H_bstride,
                                      const double* Y,
                                      const double* R, int64_t R_bstride, int64_t R_tstride,
                                      double jitter,
                                      const double* A_smooth, const double* Q_smooth,
                                      const double* dt_smooth, int64_t dt_smooth_bstride,
                                      const double* Hout, int32_t mo,
                                      void* ws, int64_t ws_bytes,
                                      double* lml, double* lml_k, double* ms, double* Ps);

/* Collocation (EKF) Kalman filter over B independent series: the state-space prior constrained by a point-wise
 * ODE / PDE residual at every step.  Replaces kf_predict_step(PDE, 'sequential') inside filter('sequential')
 * (computation/filters/kalman_filter.py:340-427, 439-485).  Per step: LTI predict; residual f = g(m_) and Jacobian
 * H_jac = dg/dx(m_) at the PREDICTED mean (:378-379; the reference: jax.jacfwd of PDE.forward_g,
 * transforms/pdes.py:236-245); optional boundary update with H, R * 0 and y = boundary[k] (:382-391); pseudo-
 * observation update with H_jac, ZERO noise, y = y_pseudo and innovation f (:395-414); data update with H, R, Y
 * when observe_data != 0 (:417-421).  lml_k is the term of the LAST update of the step, as in the reference.
 * The smoother of this model is physs_rts_smooth_f64 on (mf, Pf) (rts_smoother.py:108-150).
 *
 * Residual table (HOST pointers, read during the call; pc = 1 or 2 outputs -- up to 3 with three latents --, n_terms <= 12):
 *     g_p(x, k) = sum_j res_w[p, j] x[j] + sum_{q : term_out[q] = p} term_coef[q] * phi_{term_kind[q]}(x[term_idx[q]])
 *                 + forcing[p, k]
 *   phi: PHYSS_RES_SIN / _COS / _SQUARE / _CUBE, or the bilinear PHYSS_RES_PROD: term_coef[q] x[i] x[j] with
 *   term_idx[q] = i | (j << 8) (the x y couplings of LotkaVolterra / LorenzSystem, transforms/pdes.py:818-1090).
 *   forcing [pc, T] (DEVICE, shared by the batch) or NULL.
 *   y_pseudo [pc] (HOST): the pseudo observation, NaN = output not collocated (PDE.psuedo_observations).
 *   boundary [B, T, m] in the step layout (DEVICE, NaN = no boundary observation at that step) or NULL.
 * Shapes supported: one latent, 2 <= d <= 4 with one Matern block (DISC_MATERN, nblk = 1) or DISC_GIVEN, m = 1 or
 * identity H with m = d; systems of ODEs over 2 / 3 independent latents of state dim 2 (d = 4 / 6: DISC_MATERN with
 * nblk = d / 2 Matern-3/2 blocks, or DISC_GIVEN with dense transitions), one observation per latent (m = d / 2, any H) or
 * identity H with m = d.  Everything else as physs_kf_filter_f64. */
#define PHYSS_RES_SIN 0
#define PHYSS_RES_COS 1
#define PHYSS_RES_SQUARE 2
#define PHYSS_RES_CUBE 3
#define PHYSS_RES_PROD 4   /* bilinear: coef * x[i] * x[j], term_idx = i | (j << 8) */
int physs_kf_filter_colloc_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                               int32_t d, int32_t m,
                               int32_t disc_mode, int32_t nblk,
                               const double* A, int64_t A_bstride,
                               const double* Q, int64_t Q_bstride,
                               const double* lam, int64_t lam_bstride,
                               const double* dt, int64_t dt_bstride,
                               const double* Pinf, int64_t Pinf_bstride,
                               const double* m0, int64_t m0_bstride,
                               const double* P0, int64_t P0_bstride,
                               const double* H, int64_t H_bstride,
                               const double* Y,
                               const double* R, int64_t R_bstride, int64_t R_tstride,
                               double jitter,
                               int32_t pc, const double* res_w, int32_t n_terms, const int32_t* term_out,
                               const int32_t* term_kind, const int32_t* term_idx, const double* term_coef,
                               const double* forcing, const double* y_pseudo, const double* boundary,
                               int32_t observe_data,
                               double* mf, double* Pf, double* lml, double* lml_k);

/* Reverse pass (vector-Jacobian product) of physs_kf_filter_f64's lml: d lml[b] / d (inputs), scaled by g_lml[b].
 * Replaces `jax.jacrev` / `jax.grad` THROUGH filter('sequential') in the reference's hyper-parameter steps
 * (trainers/trainer.py:43,128-136; trainers/standard.py:58-91): the backward half of a `jax.custom_vjp` around
 * the filter call.  Same inputs as physs_kf_filter_f64 plus its outputs (mf, Pf); every matrix
 * entry counts as an independent variable (what autodiff of the reference's formulation gives, K from
 * solve(S, M H P_)).  Supported (physs_kf_vjp_supported): d <= 4, m == 1 with either discretisation (register
 * kernel, csrc/physs_vjp.cu), and d <= 32, any m <= d with PHYSS_DISC_GIVEN (lane-group kernel,
 * csrc/physs_vjp_grp.cu: full-state sites m = d at d = 6 .. 12 are the shapes a VB_NG_ADAM epoch of the
 * reference differentiates through; the chain from gA, gQ to kernel hyper-parameters is T-independent).
 *   g_lml  [B] or NULL (= 1)
 *   DISC_GIVEN : gA, gQ [B, T, d, d] in the step layout (required)
 *   DISC_MATERN: glam [B, nblk], gPinf [B, d, d] (required): the chain through A_k = expm(F(lam) dt_k) (closed
 *                forms) and Q_k = Pinf - A_k Pinf A_k^T (kernels/kernel.py:207-209) is folded on chip
 *   gH [B, m, d], gR_step [B, T, m, m] (step layout), gR_sum [B, m, m], gm0 [B, d], gP0 [B, d, d]: NULL = skip
 */
int physs_kf_vjp_supported(int32_t d, int32_t m, int32_t disc_mode, int32_t nblk);
int physs_kf_filter_vjp_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                            int32_t d, int32_t m,
                            int32_t disc_mode, int32_t nblk,
                            const double* A, int64_t A_bstride,
                            const double* Q, int64_t Q_bstride,
                            const double* lam, int64_t lam_bstride,
                            const double* dt, int64_t dt_bstride,
                            const double* Pinf, int64_t Pinf_bstride,
                            const double* m0, int64_t m0_bstride,
                            const double* P0, int64_t P0_bstride,
                            const double* H, int64_t H_bstride,
                            const double* Y,
                            const double* R, int64_t R_bstride, int64_t R_tstride,
                            double jitter,
                            const double* mf, const double* Pf, const double* g_lml,
                            double* gA, double* gQ, double* glam, double* gPinf, double* gH,
                            double* gR_step, double* gR_sum, double* gm0, double* gP0);

/* ------------------------------------------------------------------------------------------------
 * Parallel-in-time forms.  Replace filter('parallel') / smoother('parallel')
 * (computation/filters/parallel_kalman_filter.py:225-336 with the elements :73-175 and the operator
 * :178-220; parallel_rts_smoother.py:57-103 with :21-55).  Same inputs / outputs as the sequential entry
 * points; the time axis is cut into chunks of `chunk_len` steps, each chunk is folded into ONE scan
 * element on chip, the reference's associative operators run between chunks only, and all chunks are then
 * replayed concurrently by the sequential kernels (physs_gp_b200/csrc/physs_pscan.cu).
 *
 * Result contract: the SEQUENTIAL reference result (the default filter_type; SURVEY.md quirk Q1 -- the
 * reference's own parallel path starts from 2 P_inf and is inconsistent with its sequential path).
 * With jitter == 0 the scan is exact.  With jitter != 0 the reference's sequential recursion (jitter in
 * the gain solve only, kalman_filter.py:144-211 + linalg.py:29-33) is not representable by scan elements;
 * `polish` fix-up passes re-run every chunk from the previous chunk's replayed end state until it agrees
 * with the stored result to `delta` (relative) for `patience` consecutive steps.  *status (device int,
 * may be NULL) is 1 if some chunk reached its end without agreeing (raise chunk_len or polish).
 *
 *   ws: device workspace of physs_pscan_workspace_bytes(B, T, d, chunk_len) bytes, 16-byte aligned; it
 *       carries state from *_local to *_finish.  lml_k may be NULL (the workspace is used).
 *
 * Time-sharded use (one long series over several GPUs): every rank calls *_local
 * on its own time range and receives the scan element of the whole range (`total`: filter
 * [B, 3 d^2 + 2 d] = [A | C | J | b | eta], smoother [B, 2 d^2 + d] = [E | L | g]); the totals are
 * all-gathered (NCCL); *_fold pushes the prior through the totals of the preceding ranks (filter) or the
 * terminal state through those of the following ranks (smoother); *_finish replays the local range from
 * that carried state.
 */
int64_t physs_pscan_workspace_bytes(int64_t B, int64_t T, int32_t d, int64_t chunk_len);

int physs_pscan_filter_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                           int32_t d, int32_t m, int32_t disc_mode, int32_t nblk,
                           const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                           const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                           const double* Pinf, int64_t Pinf_bstride, const double* m0, int64_t m0_bstride,
                           const double* P0, int64_t P0_bstride, const double* H, int64_t H_bstride,
                           const double* Y, const double* R, int64_t R_bstride, int64_t R_tstride,
                           double jitter,
                           int64_t chunk_len, int32_t polish, double delta, int32_t patience, void* ws,
                           double* mf, double* Pf, double* lml, double* lml_k, int32_t* status);

/* Speculative parallel-in-time variants (no reference counterpart; an optimisation of the above for priors
 * whose filter forgets its initial state within `warm` <= chunk_len steps, e.g. every stationary Matern
 * model): no summaries and no scan -- chunk c > 0 of the filter starts `warm` steps early from (m0, P0),
 * chunk c of the smoother `warm` steps late from the filtered state there, and the same fix-up passes
 * (polish >= 1) both VERIFY each chunk against a restart from its neighbour's replayed state and repair it.
 * *status = 1 means some chunk still disagreed in the last pass: rerun with the exact scan entry points.
 * The smoother variant needs full-state output (Hout = NULL). */
int physs_pscan_filter_spec_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                                int32_t d, int32_t m, int32_t disc_mode, int32_t nblk,
                                const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                                const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                                const double* Pinf, int64_t Pinf_bstride, const double* m0, int64_t m0_bstride,
                                const double* P0, int64_t P0_bstride, const double* H, int64_t H_bstride,
                                const double* Y, const double* R, int64_t R_bstride, int64_t R_tstride,
                                double jitter,
                                int64_t chunk_len, int64_t warm, int32_t polish, double delta, int32_t patience,
                                void* ws, double* mf, double* Pf, double* lml, double* lml_k, int32_t* status);

int physs_pscan_smooth_spec_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                                int32_t d, int32_t disc_mode, int32_t nblk,
                                const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                                const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                                const double* Pinf, int64_t Pinf_bstride,
                                const double* mf, const double* Pf, const double* Hout, int32_t mo, double jitter,
                                int64_t chunk_len, int64_t warm, int32_t polish, double delta, int32_t patience,
                                void* ws, double* ms, double* Ps, int32_t* status);

int physs_pscan_filter_local_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                                 int32_t d, int32_t m, int32_t disc_mode, int32_t nblk,
                                 const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                                 const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                                 const double* Pinf, int64_t Pinf_bstride, const double* m0, int64_t m0_bstride,
                                 const double* P0, int64_t P0_bstride, const double* H, int64_t H_bstride,
                                 const double* Y, const double* R, int64_t R_bstride, int64_t R_tstride,
                                 double jitter,
                                 int64_t chunk_len, void* ws, double* total);

/* start_m [B, d], start_P [B, d, d]: filtered state just before this time range (NULL = (m0, P0)). */
int physs_pscan_filter_finish_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                                  int32_t d, int32_t m, int32_t disc_mode, int32_t nblk,
                                  const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                                  const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                                  const double* Pinf, int64_t Pinf_bstride, const double* m0, int64_t m0_bstride,
                                  const double* P0, int64_t P0_bstride, const double* H, int64_t H_bstride,
                                  const double* Y, const double* R, int64_t R_bstride, int64_t R_tstride,
                                  double jitter,
                                  int64_t chunk_len, int32_t polish, double delta, int32_t patience, void* ws,
                                  const double* start_m, const double* start_P,
                                  double* mf, double* Pf, double* lml, double* lml_k, int32_t* status);

/* (m_out, P_out) = (m0, P0) pushed through totals[0 .. K-1] ([K, B, 3 d^2 + 2 d], time order). */
int physs_pscan_filter_fold_f64(void* stream, int64_t B, int32_t d, int64_t K, const double* totals,
                                const double* m0, int64_t m0_bstride, const double* P0, int64_t P0_bstride,
                                double* m_out, double* P_out);

int physs_pscan_smooth_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                           int32_t d, int32_t disc_mode, int32_t nblk,
                           const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                           const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                           const double* Pinf, int64_t Pinf_bstride,
                           const double* mf, const double* Pf, const double* Hout, int32_t mo, double jitter,
                           int64_t chunk_len, void* ws, double* ms, double* Ps);

int physs_pscan_smooth_local_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                                 int32_t d, int32_t disc_mode, int32_t nblk,
                                 const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                                 const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                                 const double* Pinf, int64_t Pinf_bstride,
                                 const double* mf, const double* Pf, const double* Hout, int32_t mo, double jitter,
                                 int64_t chunk_len, void* ws, double* total);

/* start_m [B, d], start_P [B, d, d]: smoothed state of the first step AFTER this time range (NULL = this
 * range ends the series: terminal condition smoothed = filtered at T - 1).  On a range that does not end
 * the series, dt[T - 1] must be the gap to that next step. */
int physs_pscan_smooth_finish_f64(void* stream, int64_t B, int64_t T, int64_t step_bstride, int64_t step_tstride,
                                  int32_t d, int32_t disc_mode, int32_t nblk,
                                  const double* A, int64_t A_bstride, const double* Q, int64_t Q_bstride,
                                  const double* lam, int64_t lam_bstride, const double* dt, int64_t dt_bstride,
                                  const double* Pinf, int64_t Pinf_bstride,
                                  const double* mf, const double* Pf, const double* Hout, int32_t mo, double jitter,
                                  int64_t chunk_len, void* ws, const double* start_m, const double* start_P,
                                  double* ms, double* Ps);

/* (m_out, P_out) = the state (m_end, P_end) [B, d], [B, d, d] pulled back through totals[K-1 .. 0]
 * ([K, B, 2 d^2 + d], time order). */
int physs_pscan_smooth_fold_f64(void* stream, int64_t B, int32_t d, int64_t K, const double* totals,
                                const double* m_end, const double* P_end, double* m_out, double* P_out);

/* CVI likelihood kinds */
#define PHYSS_LIK_GAUSS 0            /* y = W u + e, e ~ N(0, noise): closed-form block ELL               */
#define PHYSS_LIK_POISSON_EXP 1      /* independent Poisson(binsize * exp(f_p)), f = W u, Gauss-Hermite   */
#define PHYSS_LIK_BERNOULLI_PROBIT 2 /* independent Bernoulli(Phi(f_p)) with the reference's +1e-5 jitter */
#define PHYSS_LIK_GIVEN 3            /* dELL/dm, dELL/dS supplied by the caller (e.g. jax.grad of an MC ELL) */

/* One CVI natural-gradient site update for N independent site blocks (N = B * T).
 * Replaces natural_gradients(FullConjugateGaussian) (computation/natural_gradients/cvi_nat_grad.py:
 * 346-410) + cvi_block_update (:47-87) + theta_to_lambda / lambda_to_theta
 * (exponential_family_transforms.py:25-42,70-83) + the 'NG_Moment' re-entry
 * (cvi_parameterisations.py:63-93), with the ELL gradients in closed form / Gauss-Hermite instead of
 * jax.grad (cvi_nat_grad.py:381-383), or supplied (PHYSS_LIK_GIVEN).
 *   Ytil [N, D], Vtil [N, D, D]    sites (the surrogate model's data and BlockDiagonalGaussian variance)
 *   q_mu [N, D], q_var [N, D, D]   posterior marginals of the site blocks (surrogate.posterior_blocks())
 *   y [N, P] data (NaN = missing); W [P, D] (NULL = identity, P == D) maps the block to likelihood inputs
 *   noise [., P, P] Gaussian noise, noise_stride elements between blocks (0 = shared)
 *   lik_param: Poisson binsize;  K, ghx[K], ghw[K]: Gauss-Hermite nodes and weights / sqrt(pi)
 *   beta: step size;  ng_jitter: settings.ng_jitter
 * Outputs: Ytil_out, Vtil_out (may alias the inputs); ell_out [N] or NULL: per-block data ELL.
 */
int physs_cvi_natgrad_step_f64(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik,
                               const double* Ytil, const double* Vtil,
                               const double* q_mu, const double* q_var,
                               const double* y, const double* W,
                               const double* noise, int64_t noise_stride,
                               double lik_param, int32_t K, const double* ghx, const double* ghw,
                               const double* dm_in, const double* dS_in,
                               double beta, double ng_jitter,
                               double* Ytil_out, double* Vtil_out, double* ell_out);

/* The same update for sites stored in the PRECISION parameterisation ('NG_Precision'): Ptil [N, D, D] is the site
 * precision (the surrogate's PrecisionBlockDiagonalGaussian, likelihood/gaussian.py:96-105) in and out.  Replaces
 * theta_precision_to_lambda / lambda_to_theta_precision (exponential_family_transforms.py:44-53,85-95) around
 * cvi_block_update, i.e. natural_gradients(VGP, FullConjugateGaussian, "NG_Precision")
 * (cvi_parameterisations.py:95-113 with cvi_nat_grad_utils.py:62-63): lambda_2 = -1/2 Ptil,
 * lambda_1 = (Ptil + ng_jitter I)^-1 Ytil -- the reference's own cholesky_solve with the precision's factor --,
 * Ptil' = -2 lambda_2', Ytil' = (Ptil' + ng_jitter I)^-1 lambda_1'.  Everything else as above. */
int physs_cvi_natgrad_step_prec_f64(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik,
                                    const double* Ytil, const double* Ptil,
                                    const double* q_mu, const double* q_var,
                                    const double* y, const double* W,
                                    const double* noise, int64_t noise_stride,
                                    double lik_param, int32_t K, const double* ghx, const double* ghw,
                                    const double* dm_in, const double* dS_in,
                                    double beta, double ng_jitter,
                                    double* Ytil_out, double* Ptil_out, double* ell_out);

/* Per-block expected log-likelihood (and optionally its gradients) under q = N(q_mu, q_var).
 * Replaces full_gaussian_expected_log_likelihood (computation/elbos/expected_log_likelihoods.py:90-117)
 * vmapped over blocks (dispatched_ell.py:47-132) and the non-Gaussian approximate_expectation route
 * (dispatched_ell.py:406-434) with Gauss-Hermite quadrature.  The CVI ELBO (elbos.py:163-194) is
 *   sum(ell_data) - sum(ell_surrogate) + lml_surrogate,
 * where ell_surrogate is this call with lik = GAUSS, W = NULL, y = Ytil, noise = Vtil.
 */
int physs_cvi_ell_f64(void* stream, int64_t N, int32_t D, int32_t P, int32_t lik,
                      const double* q_mu, const double* q_var,
                      const double* y, const double* W,
                      const double* noise, int64_t noise_stride,
                      double lik_param, int32_t K, const double* ghx, const double* ghw,
                      double* ell_out, double* dm_out, double* dS_out);

/* Expected log-likelihood, its mean gradient and the (Gauss-Newton) site curvature of the PHYSS-GP
 * damped-oscillator model: outputs T(u) = [x, x_tt + (g/l) sin x + b x_t] of the derivative-augmented block
 * u (x, x_t, x_tt at block positions i0, i1, i2), Gaussian noise var_obs on the observation of x and
 * var_col on the collocation residual.  Replaces, for this transform, MultiOutput([OutputMap,
 * DampedPendulum1D]).forward (transforms/pdes.py:530-597) + the Monte-Carlo ELL (integrals/approximators.py:
 * 16-58) + its jax.grad (natural_gradients/cvi_nat_grad.py:381-383) + the Laplace-Gauss-Newton delta-u
 * curvature (cvi_hessian_approximations.py:333-431,542-577).  The expectation over sin(x) is closed-form.
 *   y [N, 2]: (observation of x, collocation target, normally 0); NaN = absent at that step
 *   gauss_newton != 0: dS = -1/2 sum_p J_p^T J_p / var_p with J at u = q_mu (enforce_psd_type=
 *   'laplace_gauss_newton_delta_u'); == 0: the exact dELL/dS.
 * Outputs (each may be NULL): ell [N], dm [N, D], dS [N, D, D]; feed dm, dS to physs_cvi_natgrad_step_f64
 * with PHYSS_LIK_GIVEN.  The linear diffusion residual f_t - f_xx needs no special kernel: it is
 * PHYSS_LIK_GAUSS with W = the residual's row vector. */
int physs_cvi_ell_pendulum_f64(void* stream, int64_t N, int32_t D, int32_t i0, int32_t i1, int32_t i2,
                               const double* q_mu, const double* q_var, const double* y,
                               double g_over_l, double damping, double var_obs, double var_col,
                               int32_t gauss_newton, double* ell_out, double* dm_out, double* dS_out);

/* Gauss-Newton site curvature from caller-supplied Jacobians, for ANY prior transform: dS = -1/2 sum_p mask_p
 * J_p^T J_p / var_p per site block (cvi_hessian_approximations.py:380-431,483-486,574; the Jacobians
 * J [N, P, D] = d T(u)/du are what the reference computes with jax.jacfwd, :358-368, and stay on the JAX side).
 *   var [., P] conditional likelihood variances (var_stride elements between blocks, 0 = shared)
 *   y [N, P] or NULL: NaN entries drop the corresponding output from the sum. */
int physs_cvi_gauss_newton_f64(void* stream, int64_t N, int32_t D, int32_t P, const double* J, const double* var,
                               int64_t var_stride, const double* y, double* dS_out);

/* Batched inverse of N small SPD matrices, out[n] = (A[n] + jitter I)^-1 by Cholesky (D <= 8; non-PD -> NaN).
 * Turns precision-parameterised sites R_inv [., T, m, m] into the covariance form the filter entry points take:
 * the reference's sequential filter has no precision path (kf_update_step_with_lik_precision,
 * kalman_filter.py:43-126, raises), its parallel filter factors R_inv the same way
 * (parallel_kalman_filter.py:34-71,117-141). */
int physs_spd_inverse_f64(void* stream, int64_t N, int32_t D, const double* A, double jitter, double* out);

/* ---- Separable spatio-temporal prior, ONE series, large state (BASELINE config 2: d = Ns * ds = 400, m = Ns = 200).
 * Replaces kf_predict_step / kf_update_step / rts_smoother_step (kalman_filter.py:144-241,439-485;
 * rts_smoother.py:48-106,162-192) for the prior of kernels/kernel.py:213-265 + ss_utils.py:41-53:
 *   A_k = I_Ns (x) At_k,  Q_k = Ks (x) Qt_k  (Qt_k = Pinf_t - At_k Pinf_t At_k^T, kernel.py:207-209),
 *   H = I_Ns (x) [1 0 .. 0]  (state index = s * ds + j: the observation picks the first temporal state of every
 *   spatial point).  The Kronecker structure is used (predict O(d^2) instead of 2 d^3); Cholesky, triangular
 *   solves and the rank-2m covariance update are hand-written tile kernels on DMMA.8x8x4 inside persistent
 *   cooperative kernels -- no cuBLAS / cuSOLVER.
 *   At, Qt [nA, ds, ds]  temporal transition / process noise per DISTINCT step size, 1 <= ds <= 4
 *   idx    [T] DEVICE int32: entry k selects the (At, Qt) pair of step k.  Filter: dt[0] = 0 (At = I, Qt = 0),
 *          dt[k] = t_k - t_{k-1}; smoother: dt[k] = t_{k+1} - t_k (entry T-1 unused).
 *   Ks [Ns, Ns], m0 [d], P0 [d, d], Y [T, Ns] (NaN = missing), R [., Ns, Ns] with R_tstride elements between
 *   steps (0 = shared); ws: 16-byte aligned device workspace of physs_kron_workspace_bytes(T, Ns, ds, smoother).
 * Outputs: mf [T, d], Pf [T, d, d], lml [1];  smoother: project = 0 -> ms [T, d], Ps [T, d, d];
 * project = 1 -> H ms [T, Ns], H Ps H^T [T, Ns, Ns].  Stream-ordered, NaN on numerical failure. */
int64_t physs_kron_workspace_bytes(int64_t T, int32_t Ns, int32_t ds, int32_t smoother);
/* measurement aid: with PHYSS_KRON_PROF=1 in the environment CTA 0 accumulates nanoseconds per phase into 32
 * doubles at this offset (in doubles) of the workspace */
int64_t physs_kron_prof_offset(int64_t T, int32_t Ns, int32_t ds, int32_t smoother);
int physs_kf_filter_kron_f64(void* stream, int64_t T, int32_t Ns, int32_t ds, const double* At, const double* Qt,
                             const int32_t* idx, const double* Ks, const double* m0, const double* P0,
                             const double* Y, const double* R, int64_t R_tstride, double jitter, void* ws,
                             int64_t ws_bytes, double* mf, double* Pf, double* lml);
int physs_rts_smooth_kron_f64(void* stream, int64_t T, int32_t Ns, int32_t ds, const double* At, const double* Qt,
                              const int32_t* idx, const double* Ks, const double* mf, const double* Pf,
                              int32_t project, double jitter, void* ws, int64_t ws_bytes, double* ms, double* Ps);

/* ---- CVI site update and surrogate ELL for LARGE site blocks (D <= 208: the D = Ns = 200 blocks of config 2, one block
 * per time step), one CTA per block on the shared-memory Cholesky of the separable-prior kernels above.  Same algebra as
 * physs_cvi_natgrad_step_f64 (cvi_nat_grad.py:47-87, exponential_family_transforms.py:25-95) with the ELL gradients
 * supplied (LIK_GIVEN): dm [T, D]; dS [T, D, D], or its diagonal [T, D] with dS_diag = 1 (Gaussian likelihood with
 * diagonal noise: dm = (y - m) / s2, dS = -1 / (2 s2), zero at missing entries).  Yt [T, D], Vt [T, D, D] sites in,
 * Yn / Vn out (may not alias the inputs); qm [T, D], qS [T, D, D] posterior marginals of the blocks.
 * physs_cvi_ell_sur_big_f64: ell[t] = log N(Yt_t | qm_t, Vt_t) - 1/2 tr(Vt_t^-1 qS_t) (expected_log_likelihoods.py:90-117;
 * sites carry no missing entries).  ws: 16-byte aligned, physs_cvi_big_workspace_bytes(D) bytes. */
int64_t physs_cvi_big_workspace_bytes(int32_t D);
int physs_cvi_natgrad_big_f64(void* stream, int64_t T, int32_t D, const double* Yt, const double* Vt, const double* qm,
                              const double* dm, const double* dS, int32_t dS_diag, double beta, double ngj, void* ws,
                              int64_t ws_bytes, double* Yn, double* Vn);
int physs_cvi_ell_sur_big_f64(void* stream, int64_t T, int32_t D, const double* Yt, const double* Vt, const double* qm,
                              const double* qS, void* ws, int64_t ws_bytes, double* ell);

/* ---- Spatial conditional after the smoother (second half): posterior at N new spatial points from the
 * smoothed posterior (m_t, P_t) at the M inducing points, every time step independently.  Replaces the vmapped
 * `gaussian_spatial_conditional_cholesky` (computation/marginals.py:82-113) that `spatial_conditional_block`
 * (computation/spatial_conditionals.py:30-207, the f_only branch :150-207) runs per time step:
 *     mu_t  = W m_t                                   W  = Ksz Kzz^-1                  [N, M]
 *     var_t = ktt_t C0 + W (P_t + jitter I) W^T       C0 = Kss - Ksz Kzz^-1 Kzs        [N, N]
 * (the reference factors P_t + jitter I, :137-141, and multiplies the factor back, marginals.py:104-107).  W and C0 are
 * time-invariant and prepared by the caller; ktt [T] = the temporal kernel's variance at each step (`Ktt`, :84) or NULL
 * for 1.  m [T, M], P [T, M, M] (symmetric), mu [T, N]; var [T, N, N] (exactly symmetric) or, with diagonal = 1, only
 * its diagonal [T, N].  One persistent CTA per SM, tile GEMMs on DMMA.  ws: 16-byte aligned,
 * physs_spatial_conditional_ws_bytes(M, N) bytes. */
int64_t physs_spatial_conditional_ws_bytes(int32_t M, int32_t N);
int physs_spatial_conditional_f64(void* stream, int64_t T, int32_t M, int32_t N, const double* W, const double* C0,
                                  const double* ktt, const double* m, const double* P, double jitter,
                                  int32_t diagonal, void* ws, int64_t ws_bytes, double* mu, double* var);

/* Per-series sums over the time axis: out[b] = sum_k (x[b, k] - sub[b, k]) for a [B, T] array with element strides
 * (bstride, tstride) -- either step layout; sub may be NULL.  The ELL sums of `elbo` (elbos.py:163-194: `np.sum` of the
 * per-block expected log-likelihoods, data minus surrogate) in two deterministic stages, coalesced in the time-major
 * layout.  scratch: B * ceil(T / 512) doubles. */
int physs_sum_steps_f64(void* stream, int64_t B, int64_t T, int64_t bstride, int64_t tstride, const double* x,
                        const double* sub, double* scratch, double* out);

/* FP64 FMA throughput probe (measurement aid for the FP64-pipe roofline; no reference counterpart).
 * Launches blocks x 256 threads doing iters x 8 independent FMAs each: flops = blocks*256*iters*16. */
int physs_fp64_probe(void* stream, int32_t blocks, int64_t iters, double* out);

#ifdef __cplusplus
}
#endif
#endif /* PHYSS_B200_H_ */
