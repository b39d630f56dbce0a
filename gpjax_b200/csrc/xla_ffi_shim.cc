// XLA FFI handlers over the C ABI (include/gpjax_b200.h) -- the `jax.ffi` binding the north-star names.
//
// NOT compiled in this image: jax / jaxlib (and therefore `xla/ffi/api/ffi.h`, shipped in
// jax.ffi.include_dir()) are not installed, so the whole translation unit is guarded.  On a machine with
// jax:   g++ -O2 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//            xla_ffi_shim.cc -L../lib -lgpjax_b200 -o libgpjax_b200_xla.so
// Every handler only forwards XLA-owned buffers and the execution stream to a gpb_* entry point:
// outputs are pre-allocated by XLA, scratch comes from ffi::ScratchAllocator, errors become ffi::Error,
// nothing synchronises.  See INTEGRATION.md for the Python side (register_ffi_target + custom_vjp).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define GPB_HAVE_XLA_FFI 1
#endif
#endif

#ifdef GPB_HAVE_XLA_FFI
#include <cuda_runtime_api.h>
#include "xla/ffi/api/ffi.h"
#include "../../include/gpjax_b200.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;

static ffi::Error to_error(int rc, const char* what) {
    if (rc == GPB_OK) return ffi::Error::Success();
    return ffi::Error(rc == GPB_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented : ffi::ErrorCode::kInvalidArgument,
                      std::string(what) + " failed with GPB error " + std::to_string(rc));
}

// K = gram(X, Z)   (kernel.gram / kernel.cross_covariance, gpjax/kernels/computations/dense.py:32-36)
static ffi::Error GramImpl(cudaStream_t stream, F64 X, F64 Z, F64 ell, F64 var, ffi::Result<F64> K, int32_t kind,
                           double diag_add) {
    const int64_t N = X.dimensions()[0], D = X.dimensions()[1], M = Z.dimensions()[0];
    const int iso = ell.element_count() == 1 ? 1 : 0;
    return to_error(gpb_gram(stream, kind, N, M, (int)D, X.typed_data(), D, Z.typed_data(), D, ell.typed_data(), iso,
                             var.typed_data(), diag_add, nullptr, 0, K->typed_data(), M),
                    "gpb_gram");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbGram, GramImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<double>("diag_add"));

// conjugate_mll forward: value, alpha, and the Sigma/L buffer + workspace kept as residuals for the VJP
static ffi::Error MllFwdImpl(cudaStream_t stream, F64 X, F64 y, F64 ell, F64 var, F64 sn, F64 mean,
                             ffi::Result<F64> value, ffi::Result<F64> alpha, ffi::Result<F64> sigma,
                             ffi::Result<F64> ws, ffi::Result<S32> info, int32_t kind, double jitter) {
    const int64_t N = X.dimensions()[0], D = X.dimensions()[1];
    const int iso = ell.element_count() == 1 ? 1 : 0;
    cudaMemsetAsync(info->typed_data(), 0, sizeof(int32_t), stream);
    return to_error(gpb_mll_forward(stream, kind, N, (int)D, X.typed_data(), D, y.typed_data(), ell.typed_data(), iso,
                                    var.typed_data(), sn.typed_data(), mean.typed_data(), jitter, sigma->typed_data(), N,
                                    ws->typed_data(), (int64_t)ws->size_bytes(), value->typed_data(),
                                    alpha->typed_data(), info->typed_data()),
                    "gpb_mll_forward");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbMllFwd, MllFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<S32>()
                                  .Attr<int32_t>("kind").Attr<double>("jitter"));

// conjugate_mll backward: sigma / ws are donated (input_output_aliases) because potri works in place
static ffi::Error MllBwdImpl(cudaStream_t stream, F64 X, F64 ell, F64 var, F64 sn, F64 sigma, F64 ws, F64 alpha,
                             F64 gout, ffi::Result<F64> g_ell, ffi::Result<F64> g_var, ffi::Result<F64> g_sn,
                             ffi::Result<F64> g_mean, int32_t kind) {
    const int64_t N = X.dimensions()[0], D = X.dimensions()[1];
    const int iso = ell.element_count() == 1 ? 1 : 0;
    return to_error(gpb_mll_backward(stream, kind, N, (int)D, X.typed_data(), D, ell.typed_data(), iso, var.typed_data(),
                                     sn.typed_data(), sigma.typed_data(), N, ws.typed_data(), (int64_t)ws.size_bytes(),
                                     alpha.typed_data(), gout.typed_data(), g_ell->typed_data(), g_var->typed_data(),
                                     g_sn->typed_data(), g_mean->typed_data()),
                    "gpb_mll_backward");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbMllBwd, MllBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("kind"));
#endif  // GPB_HAVE_XLA_FFI
