"""ctypes binding of libphyss_b200.so (the C ABI declared in include/physs_b200.h).

There is deliberately NO fallback: if the CUDA library has not been built, importing the compute
entry points raises.  Build it with `python -m physs_gp_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libphyss_b200.so")

PHYSS_OK = 0
DISC_GIVEN = 0
DISC_MATERN = 1
DISC_IWP = 2
ABI_VERSION = 14

_c_i32 = ctypes.c_int32
_c_i64 = ctypes.c_int64
_c_f64 = ctypes.c_double
_ptr = ctypes.c_void_p

# shared argument prefixes of the filter / smoother families (include/physs_b200.h)
_FILTER_HEAD = [_ptr, _c_i64, _c_i64, _c_i64, _c_i64, _c_i32, _c_i32, _c_i32, _c_i32,
                _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64,
                _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64,
                _ptr, _ptr, _c_i64, _c_i64, _c_f64]
_SMOOTH_HEAD = [_ptr, _c_i64, _c_i64, _c_i64, _c_i64, _c_i32, _c_i32, _c_i32,
                _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64,
                _ptr, _ptr, _ptr, _c_i32, _c_f64]

# name -> (restype, argtypes); must list every symbol include/physs_b200.h declares.
SIGNATURES = {
    "physs_abi_version": (ctypes.c_int, []),
    "physs_last_error": (ctypes.c_char_p, []),
    "physs_kf_supported": (ctypes.c_int, [_c_i32, _c_i32, _c_i32, _c_i32]),
    "physs_kf_wave_series": (_c_i64, [_c_i32, _c_i32, _c_i32]),
    "physs_kf_filter_f64": (ctypes.c_int, [
        _ptr, _c_i64, _c_i64, _c_i64, _c_i64, _c_i32, _c_i32, _c_i32, _c_i32,
        _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64,
        _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64,
        _ptr, _ptr, _c_i64, _c_i64, _c_f64,
        _ptr, _ptr, _ptr, _ptr]),
    "physs_rts_smooth_f64": (ctypes.c_int, [
        _ptr, _c_i64, _c_i64, _c_i64, _c_i64, _c_i32, _c_i32, _c_i32,
        _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64, _ptr, _c_i64,
        _ptr, _ptr, _ptr, _c_i32, _c_f64, _ptr, _ptr]),
    "physs_kf_filter_smooth_f64": (ctypes.c_int, _FILTER_HEAD + [_ptr, _ptr, _ptr, _c_i64, _ptr, _c_i32,
                                                                _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "physs_kf_filter_smooth_packed_supported": (ctypes.c_int, [_c_i32, _c_i32, _c_i32, _c_i32]),
    "physs_kf_filter_smooth_packed_ws_bytes": (_c_i64, [_c_i64, _c_i64, _c_i64, _c_i32]),
    "physs_kf_filter_smooth_packed_f64": (ctypes.c_int, _FILTER_HEAD + [_ptr, _ptr, _ptr, _c_i64, _ptr, _c_i32,
                                                                       _ptr, _c_i64, _ptr, _ptr, _ptr, _ptr]),
    "physs_kf_filter_colloc_f64": (ctypes.c_int, _FILTER_HEAD + [_c_i32, _ptr, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                                                _ptr, _c_i32, _ptr, _ptr, _ptr, _ptr]),
    "physs_kf_vjp_supported": (ctypes.c_int, [_c_i32, _c_i32, _c_i32, _c_i32]),
    "physs_kf_filter_vjp_f64": (ctypes.c_int, _FILTER_HEAD + [_ptr] * 12),
    "physs_pscan_workspace_bytes": (_c_i64, [_c_i64, _c_i64, _c_i32, _c_i64]),
    "physs_pscan_filter_f64": (ctypes.c_int, _FILTER_HEAD + [_c_i64, _c_i32, _c_f64, _c_i32, _ptr,
                                                            _ptr, _ptr, _ptr, _ptr, _ptr]),
    "physs_pscan_filter_spec_f64": (ctypes.c_int, _FILTER_HEAD + [_c_i64, _c_i64, _c_i32, _c_f64, _c_i32, _ptr,
                                                                 _ptr, _ptr, _ptr, _ptr, _ptr]),
    "physs_pscan_smooth_spec_f64": (ctypes.c_int, _SMOOTH_HEAD + [_c_i64, _c_i64, _c_i32, _c_f64, _c_i32, _ptr,
                                                                 _ptr, _ptr, _ptr]),
    "physs_pscan_filter_local_f64": (ctypes.c_int, _FILTER_HEAD + [_c_i64, _ptr, _ptr]),
    "physs_pscan_filter_finish_f64": (ctypes.c_int, _FILTER_HEAD + [_c_i64, _c_i32, _c_f64, _c_i32, _ptr, _ptr, _ptr,
                                                                   _ptr, _ptr, _ptr, _ptr, _ptr]),
    "physs_pscan_filter_fold_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i64, _ptr, _ptr, _c_i64, _ptr, _c_i64,
                                                   _ptr, _ptr]),
    "physs_pscan_smooth_f64": (ctypes.c_int, _SMOOTH_HEAD + [_c_i64, _ptr, _ptr, _ptr]),
    "physs_pscan_smooth_local_f64": (ctypes.c_int, _SMOOTH_HEAD + [_c_i64, _ptr, _ptr]),
    "physs_pscan_smooth_finish_f64": (ctypes.c_int, _SMOOTH_HEAD + [_c_i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "physs_pscan_smooth_fold_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "physs_cvi_ell_pendulum_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr,
                                                  _c_f64, _c_f64, _c_f64, _c_f64, _c_i32, _ptr, _ptr, _ptr]),
    "physs_cvi_gauss_newton_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i32, _ptr, _ptr, _c_i64, _ptr, _ptr]),
    "physs_fp64_probe": (ctypes.c_int, [_ptr, _c_i32, _c_i64, _ptr]),
    "physs_cvi_big_workspace_bytes": (_c_i64, [_c_i32]),
    "physs_cvi_natgrad_big_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i32, _c_f64, _c_f64,
                                                 _ptr, _c_i64, _ptr, _ptr]),
    "physs_cvi_ell_sur_big_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64, _ptr]),
    "physs_sum_steps_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i64, _c_i64, _c_i64, _ptr, _ptr, _ptr, _ptr]),
    "physs_spatial_conditional_ws_bytes": (_c_i64, [_c_i32, _c_i32]),
    "physs_spatial_conditional_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _c_f64,
                                                     _c_i32, _ptr, _c_i64, _ptr, _ptr]),
    "physs_kron_workspace_bytes": (_c_i64, [_c_i64, _c_i32, _c_i32, _c_i32]),
    "physs_kron_prof_offset": (_c_i64, [_c_i64, _c_i32, _c_i32, _c_i32]),
    "physs_kf_filter_kron_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                                _ptr, _ptr, _c_i64, _c_f64, _ptr, _c_i64, _ptr, _ptr, _ptr]),
    "physs_rts_smooth_kron_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                                 _c_i32, _c_f64, _ptr, _c_i64, _ptr, _ptr]),
    "physs_spd_inverse_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _ptr, _c_f64, _ptr]),
    "physs_cvi_natgrad_step_f64": (ctypes.c_int, [
        _ptr, _c_i64, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64,
        _c_f64, _c_i32, _ptr, _ptr, _ptr, _ptr, _c_f64, _c_f64, _ptr, _ptr, _ptr]),
    "physs_cvi_natgrad_step_prec_f64": (ctypes.c_int, [
        _ptr, _c_i64, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64,
        _c_f64, _c_i32, _ptr, _ptr, _ptr, _ptr, _c_f64, _c_f64, _ptr, _ptr, _ptr]),
    "physs_cvi_ell_f64": (ctypes.c_int, [
        _ptr, _c_i64, _c_i32, _c_i32, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64,
        _c_f64, _c_i32, _ptr, _ptr, _ptr, _ptr, _ptr]),
}

LIK_GAUSS, LIK_POISSON_EXP, LIK_BERNOULLI_PROBIT, LIK_GIVEN = 0, 1, 2, 3

BIG_LIB_PATH = os.path.join(_HERE, "libphyss_b200_big.so")
_c_i32p = ctypes.POINTER(ctypes.c_int32)
# include/physs_b200_big.h
BIG_SIGNATURES = {
    "physs_big_last_error": (ctypes.c_char_p, []),
    "physs_big_workspace_bytes": (_c_i64, [_c_i32, _c_i32]),
    "physs_kf_filter_big_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _c_i32, _ptr, _ptr, _c_i32p, _ptr, _ptr, _ptr,
                                               _ptr, _ptr, _c_i64, _c_f64, _ptr, _c_i64, _ptr, _ptr, _ptr]),
    "physs_rts_smooth_big_f64": (ctypes.c_int, [_ptr, _c_i64, _c_i32, _ptr, _ptr, _c_i32p, _ptr, _ptr, _ptr, _c_i32,
                                                _c_f64, _ptr, _c_i64, _ptr, _ptr]),
}

_lib = None
_big = None


class PhyssError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach signatures.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "physs_gp_b200: %s not found -- the CUDA library is not built. Run "
            "`python -m physs_gp_b200.build`. There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    v = lib.physs_abi_version()
    if v != ABI_VERSION:
        raise ImportError("physs_gp_b200: ABI version mismatch (lib %d, binding %d); rebuild" % (v, ABI_VERSION))
    _lib = lib
    return lib


def check(status, what):
    if status != PHYSS_OK:
        msg = load().physs_last_error()
        raise PhyssError("%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))


def load_big():
    """Load libphyss_b200_big.so (large-block path; depends on cuBLAS / cuSOLVER).  Raises if missing."""
    global _big
    if _big is not None:
        return _big
    if not os.path.exists(BIG_LIB_PATH):
        raise ImportError(
            "physs_gp_b200: %s not found -- run `python -m physs_gp_b200.build`. There is no CPU fallback."
            % BIG_LIB_PATH)
    lib = ctypes.CDLL(BIG_LIB_PATH)
    for name, (res, args) in BIG_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _big = lib
    return lib


def check_big(status, what):
    if status != 0:
        msg = load_big().physs_big_last_error()
        raise PhyssError("%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))
