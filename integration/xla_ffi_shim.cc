// xla_ffi_shim.cc -- XLA FFI (jax.ffi) handlers that forward to the C ABI of libphyss_b200.so.
//
// NOT BUILT in this repository's container: the XLA FFI headers ship inside jaxlib
// (`python -c "import jax.ffi; print(jax.ffi.include_dir())"`), and jax / jaxlib are not installable here
// (no network).  physs_gp_b200/build.py compiles *.cu only, so this file is skipped; INTEGRATION.md shows
// the one-line build a maintainer of the reference runs where jaxlib exists:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include integration/xla_ffi_shim.cc \
//       -L physs_gp_b200 -lphyss_b200 -o physs_gp_b200/libphyss_b200_ffi.so
//
// It contains no arithmetic: it unpacks XLA buffers (device pointers + dimensions) and the CUDA stream XLA
// runs the computation on, and calls the plain-C entry points of include/physs_b200.h.  The reference-side
// Python that registers these targets (`jax.ffi.register_ffi_target`) and calls them (`jax.ffi.ffi_call`)
// from `@dispatch('b200') def filter / smoother` is in INTEGRATION.md section 1.
//
// Batch layout: buffers carry a leading batch axis B (1 for the reference's un-batched calls); per-step
// arrays are batch-major [B, T, ...] (step strides (T, 1)), or time-major [T, B, ...] when the attribute
// time_major != 0 (jax.vmap(..., out_axes=1)), see the header.
#include <cstdint>

#include <cuda_runtime_api.h>

#include "physs_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;

inline ffi::Error Status(int rc) {
  if (rc == PHYSS_OK) return ffi::Error::Success();
  return ffi::Error(rc == PHYSS_ERR_BAD_ARG ? ffi::ErrorCode::kInvalidArgument
                    : rc == PHYSS_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                  : ffi::ErrorCode::kInternal,
                    physs_last_error());
}

// element stride between series of an operand that is either shared ([...]) or per series ([B, ...])
inline int64_t BatchStride(const F64& x, size_t shared_rank, int64_t per_series) {
  return x.dimensions().size() > shared_rank ? per_series : 0;
}

// filter('b200'): operands (A, Q | lam), dt, P_inf, m0, P0, H, Y, R -> results mf, Pf, lml
//   disc_mode = PHYSS_DISC_GIVEN : A, Q [B|1, T, d, d]   (lam is an empty buffer)
//   disc_mode = PHYSS_DISC_MATERN: lam [B|1, nblk], P_inf [B|1, d, d] (A, Q are empty buffers)
ffi::Error KfFilterImpl(cudaStream_t stream, F64 A, F64 Q, F64 lam, F64 dt, F64 Pinf, F64 m0, F64 P0, F64 H,
                        F64 Y, F64 R, F64Out mf, F64Out Pf, F64Out lml, int32_t disc_mode, int32_t nblk,
                        int32_t time_major, int32_t h_identity, double jitter) {
  const auto ydim = Y.dimensions();                       // [B, T, m] or [T, B, m]
  if (ydim.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "Y must be rank 3");
  const int64_t B = time_major ? ydim[1] : ydim[0];
  const int64_t T = time_major ? ydim[0] : ydim[1];
  const int32_t m = static_cast<int32_t>(ydim[2]);
  const int32_t d = static_cast<int32_t>(P0.dimensions().back());
  const int64_t dd = static_cast<int64_t>(d) * d;
  const auto rdim = R.dimensions();                       // [B|1, T|1, m, m]
  const int64_t R_t = (rdim.size() >= 3 && rdim[rdim.size() - 3] > 1) ? static_cast<int64_t>(m) * m : 0;
  const int64_t R_b = (rdim.size() == 4 && rdim[0] > 1) ? (R_t ? T * m * m : static_cast<int64_t>(m) * m) : 0;
  const int rc = physs_kf_filter_f64(
      stream, B, T, time_major ? 1 : T, time_major ? B : 1, d, m, disc_mode, nblk,
      disc_mode == PHYSS_DISC_GIVEN ? A.typed_data() : nullptr, BatchStride(A, 3, T * dd),
      disc_mode == PHYSS_DISC_GIVEN ? Q.typed_data() : nullptr, BatchStride(Q, 3, T * dd),
      disc_mode == PHYSS_DISC_MATERN ? lam.typed_data() : nullptr, BatchStride(lam, 1, nblk),
      dt.typed_data(), BatchStride(dt, 1, T),
      disc_mode == PHYSS_DISC_MATERN ? Pinf.typed_data() : nullptr, BatchStride(Pinf, 2, dd),
      m0.typed_data(), BatchStride(m0, 1, d), P0.typed_data(), BatchStride(P0, 2, dd),
      h_identity ? nullptr : H.typed_data(), BatchStride(H, 2, static_cast<int64_t>(m) * d),
      Y.typed_data(), R.typed_data(), R_b, R_t, jitter,
      mf->typed_data(), Pf->typed_data(), lml->typed_data(), /*lml_k=*/nullptr);
  return Status(rc);
}

// smoother('b200'): operands (A, Q | lam), dt, P_inf, mf, Pf, Hout -> results ms, Ps
ffi::Error RtsSmoothImpl(cudaStream_t stream, F64 A, F64 Q, F64 lam, F64 dt, F64 Pinf, F64 mf, F64 Pf, F64 Hout,
                         F64Out ms, F64Out Ps, int32_t disc_mode, int32_t nblk, int32_t time_major,
                         int32_t full_state, double jitter) {
  const auto mdim = mf.dimensions();                      // [B, T, d] or [T, B, d]
  if (mdim.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "mf must be rank 3");
  const int64_t B = time_major ? mdim[1] : mdim[0];
  const int64_t T = time_major ? mdim[0] : mdim[1];
  const int32_t d = static_cast<int32_t>(mdim[2]);
  const int64_t dd = static_cast<int64_t>(d) * d;
  const int32_t mo = full_state ? 0 : static_cast<int32_t>(Hout.dimensions().front());
  const int rc = physs_rts_smooth_f64(
      stream, B, T, time_major ? 1 : T, time_major ? B : 1, d, disc_mode, nblk,
      disc_mode == PHYSS_DISC_GIVEN ? A.typed_data() : nullptr, BatchStride(A, 3, T * dd),
      disc_mode == PHYSS_DISC_GIVEN ? Q.typed_data() : nullptr, BatchStride(Q, 3, T * dd),
      disc_mode == PHYSS_DISC_MATERN ? lam.typed_data() : nullptr, BatchStride(lam, 1, nblk),
      dt.typed_data(), BatchStride(dt, 1, T),
      disc_mode == PHYSS_DISC_MATERN ? Pinf.typed_data() : nullptr, BatchStride(Pinf, 2, dd),
      mf.typed_data(), Pf.typed_data(), full_state ? nullptr : Hout.typed_data(), mo, jitter,
      ms->typed_data(), Ps->typed_data());
  return Status(rc);
}

// filter_vjp('b200'): the backward rule of the jax.custom_vjp around the filter call.
// operands: the filter's operands, its results (mf, Pf) and the cotangent g_lml [B]
// results  : DISC_GIVEN  -> gA, gQ [B, T, d, d] (glam, gPinf empty);  DISC_MATERN -> glam [B, nblk], gPinf [B, d, d]
//            (gA, gQ empty);  always gH [B, m, d], gR [B, T, m, m] (per step), gm0 [B, d], gP0 [B, d, d]
ffi::Error KfFilterVjpImpl(cudaStream_t stream, F64 A, F64 Q, F64 lam, F64 dt, F64 Pinf, F64 m0, F64 P0, F64 H,
                           F64 Y, F64 R, F64 mf, F64 Pf, F64 g_lml, F64Out gA, F64Out gQ, F64Out glam,
                           F64Out gPinf, F64Out gH, F64Out gR, F64Out gm0, F64Out gP0, int32_t disc_mode,
                           int32_t nblk, int32_t time_major, int32_t h_identity, double jitter) {
  const auto ydim = Y.dimensions();
  if (ydim.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "Y must be rank 3");
  const int64_t B = time_major ? ydim[1] : ydim[0];
  const int64_t T = time_major ? ydim[0] : ydim[1];
  const int32_t m = static_cast<int32_t>(ydim[2]);
  const int32_t d = static_cast<int32_t>(P0.dimensions().back());
  const int64_t dd = static_cast<int64_t>(d) * d;
  const auto rdim = R.dimensions();
  const int64_t R_t = (rdim.size() >= 3 && rdim[rdim.size() - 3] > 1) ? static_cast<int64_t>(m) * m : 0;
  const int64_t R_b = (rdim.size() == 4 && rdim[0] > 1) ? (R_t ? T * m * m : static_cast<int64_t>(m) * m) : 0;
  const bool given = disc_mode == PHYSS_DISC_GIVEN;
  const int rc = physs_kf_filter_vjp_f64(
      stream, B, T, time_major ? 1 : T, time_major ? B : 1, d, m, disc_mode, nblk,
      given ? A.typed_data() : nullptr, BatchStride(A, 3, T * dd), given ? Q.typed_data() : nullptr,
      BatchStride(Q, 3, T * dd), given ? nullptr : lam.typed_data(), BatchStride(lam, 1, nblk), dt.typed_data(),
      BatchStride(dt, 1, T), given ? nullptr : Pinf.typed_data(), BatchStride(Pinf, 2, dd), m0.typed_data(),
      BatchStride(m0, 1, d), P0.typed_data(), BatchStride(P0, 2, dd), h_identity ? nullptr : H.typed_data(),
      BatchStride(H, 2, static_cast<int64_t>(m) * d), Y.typed_data(), R.typed_data(), R_b, R_t, jitter,
      mf.typed_data(), Pf.typed_data(), g_lml.typed_data(), given ? gA->typed_data() : nullptr,
      given ? gQ->typed_data() : nullptr, given ? nullptr : glam->typed_data(), given ? nullptr : gPinf->typed_data(),
      gH->typed_data(), gR->typed_data(), /*gR_sum=*/nullptr, gm0->typed_data(), gP0->typed_data());
  return Status(rc);
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    PhyssKfFilterFfi, KfFilterImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()      // A, Q, lam, dt, P_inf
        .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()      // m0, P0, H, Y, R
        .Ret<F64>().Ret<F64>().Ret<F64>()                            // mf, Pf, lml
        .Attr<int32_t>("disc_mode").Attr<int32_t>("nblk").Attr<int32_t>("time_major")
        .Attr<int32_t>("h_identity").Attr<double>("jitter"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    PhyssRtsSmoothFfi, RtsSmoothImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()      // A, Q, lam, dt, P_inf
        .Arg<F64>().Arg<F64>().Arg<F64>()                            // mf, Pf, Hout
        .Ret<F64>().Ret<F64>()                                       // ms, Ps
        .Attr<int32_t>("disc_mode").Attr<int32_t>("nblk").Attr<int32_t>("time_major")
        .Attr<int32_t>("full_state").Attr<double>("jitter"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    PhyssKfFilterVjpFfi, KfFilterVjpImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()      // A, Q, lam, dt, P_inf
        .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()      // m0, P0, H, Y, R
        .Arg<F64>().Arg<F64>().Arg<F64>()                            // mf, Pf, g_lml
        .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()                 // gA, gQ, glam, gPinf
        .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()                 // gH, gR, gm0, gP0
        .Attr<int32_t>("disc_mode").Attr<int32_t>("nblk").Attr<int32_t>("time_major")
        .Attr<int32_t>("h_identity").Attr<double>("jitter"));
