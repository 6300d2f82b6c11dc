/* physs_b200_big.h -- C ABI of libphyss_b200_big.so: the large-block (d ~ 10^2 .. 10^3) sequential Kalman
 * filter / RTS smoother for ONE series, BASELINE config 2 (separable spatio-temporal prior: state d = 2 Ns,
 * observations m = Ns).  Same reference functions as physs_kf_filter_f64 / physs_rts_smooth_f64
 * (computation/filters/kalman_filter.py:144-241,439-485; rts_smoother.py:48-106,162-192), same conventions
 * (device pointers, row-major fp64, NaN = missing, NaN-on-failure, stream-ordered); the dense products and
 * factorisations are cuBLAS / cuSOLVER calls on the caller's stream, which is why this lives in its own
 * library (libphyss_b200.so itself has no library dependency).
 *
 *   A, Q        [nA, d, d]   discretisations A_k = expm(F dt_k), Q_k, one per DISTINCT step size
 *   disc_index  [T] HOST int32 array: entry k selects the (A, Q) pair of step k (NULL = k itself, nA = T).
 *               Filter: dt[0] = 0, dt[k] = t_k - t_{k-1};  smoother: dt[k] = t_{k+1} - t_k.
 *   ws          device workspace of physs_big_workspace_bytes(d, m) bytes (16-byte aligned)
 */
#ifndef PHYSS_B200_BIG_H_
#define PHYSS_B200_BIG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHYSS_BIG_OK 0
#define PHYSS_BIG_ERR_BAD_ARG 1
#define PHYSS_BIG_ERR_LIB 2
#define PHYSS_BIG_ERR_CUDA 3

const char* physs_big_last_error(void);

int64_t physs_big_workspace_bytes(int32_t d, int32_t m);

/* m0 [d], P0 [d, d], H [m, d], Y [T, m], R [., m, m] with R_tstride elements between steps (0 = shared).
 * Outputs: mf [T, d], Pf [T, d, d], lml [1] (device). */
int physs_kf_filter_big_f64(void* stream, int64_t T, int32_t d, int32_t m, const double* A, const double* Q,
                            const int32_t* disc_index, const double* m0, const double* P0, const double* H,
                            const double* Y, const double* R, int64_t R_tstride, double jitter, void* ws,
                            int64_t ws_bytes, double* mf, double* Pf, double* lml);

/* Hout [mo, d] projects the output (mo = 0 / NULL: full state).  Outputs ms [T, mo'], Ps [T, mo', mo']. */
int physs_rts_smooth_big_f64(void* stream, int64_t T, int32_t d, const double* A, const double* Q,
                             const int32_t* disc_index, const double* mf, const double* Pf, const double* Hout,
                             int32_t mo, double jitter, void* ws, int64_t ws_bytes, double* ms, double* Ps);

#ifdef __cplusplus
}
#endif
#endif /* PHYSS_B200_BIG_H_ */
