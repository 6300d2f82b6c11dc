"""Numerically load-bearing global flags; mirror of the reference's stgp/settings.py:30-67.

They are read when a b200 op is called and passed to the kernels as runtime scalars.
"""
jitter = 1e-5          # settings.py:63 -- gain solve, smoother chol(P_pred + jitter I)
ng_jitter = 1e-7       # settings.py:64 -- theta <-> lambda transforms
kalman_filter_force_symmetric = False   # settings.py:33 (only the default is supported)
parallel_kf_force_linear_solve = False  # settings.py:55 (only the default is supported)
verbose = False

# B200 memory order of batched per-step arrays (no reference counterpart: the reference has no batch
# axis).  True: batches of >= time_major_min_batch series are stored [T, B, ...] in memory (returned as
# [B, T, ...] views), which lets a warp's 32 series move one contiguous span per step.
time_major = True
time_major_min_batch = 32
# filter_and_smooth(full_state=False) on a time-major batch with state dim <= 4: one C-ABI call whose filtered
# moments stay in a packed workspace (physs_kf_filter_smooth_packed_f64) instead of two calls around full (m, P)
# outputs; same arithmetic (lml bitwise, moments to 1e-14).  False = always the two calls (A-B timing)
fused_packed = True

# parallel-in-time path (filter_type='b200_parallel'): chunk length (None = chosen from B and T) and the
# number of fix-up passes that reconcile the scan with the jittered sequential recursion (None = 4 if
# jitter != 0 else 0); see include/physs_b200.h, "Parallel-in-time forms".
pscan_chunk_len = None
pscan_polish = None

# filter_type='b200_auto': parallel in time below this batch size and above this series length
auto_parallel_max_batch = 8192
auto_parallel_min_steps = 1024

# single series with state dim above this go to the large-block path (libphyss_b200_big.so: cuBLAS / cuSOLVER
# per step); at or below it the shared-memory lane-group kernels are used
big_block_min_dim = 32
# separable spatio-temporal priors (A = I (x) A_t) on that route use the hand-written persistent kernels
# (physs_kf_filter_kron_f64 / physs_rts_smooth_kron_f64); False sends them through the library path as well
kron_kernels = True
# read the device flag of the parallel-in-time filter (one host sync) and fall back to the sequential kernels
# when a chunk did not converge; set False to keep the call fully asynchronous
pscan_check_status = True

