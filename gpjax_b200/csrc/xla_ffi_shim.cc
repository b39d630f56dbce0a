// XLA FFI handlers over the C ABI (include/gpjax_b200.h) -- the `jax.ffi` binding the north-star names.
//
// jax / jaxlib (and therefore `xla/ffi/api/ffi.h`, shipped in jax.ffi.include_dir()) are not installable in this image, so
// this translation unit is NOT part of libgpjax_b200.so.  It is kept complete and is type-checked on every CPU test run
// against a stand-in of the public header (tests/xla_stub/xla/ffi/api/ffi.h, tests/test_abi_symbols.py); it has never been
// linked against a real jaxlib.  On a machine with jax:
//     g++ -O2 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I/usr/local/cuda/include ...
//         xla_ffi_shim.cc -L../lib -lgpjax_b200 -o libgpjax_b200_xla.so
// Every handler only forwards XLA-owned buffers and the execution stream to ONE gpb_* entry point: outputs are pre-allocated
// by XLA, workspaces are ordinary outputs (they are residuals: the backward needs them exactly as the forward left them),
// errors become ffi::Error, nothing synchronises.  In-place entry points (potrf, the MLL backward, trsv / trsm) take the
// buffer they overwrite as an operand AND return it as a result: the Python side declares `input_output_aliases` so XLA
// donates the operand; if XLA nevertheless hands over distinct buffers the handler copies first, so the
// operand is never mutated behind XLA's back.  See INTEGRATION.md for the Python side (register_ffi_target + custom_vjp;
// the two all-reduces of the sharded sparse path are jax.lax.psum between the handlers).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define GPB_HAVE_XLA_FFI 1
#endif
#endif

#ifdef GPB_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <string>

#include "../../include/gpjax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
using Stream = ffi::PlatformStream<cudaStream_t>;

static ffi::Error to_error(int rc, const char* what) {
    if (rc == GPB_OK) return ffi::Error::Success();
    return ffi::Error(rc == GPB_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented : ffi::ErrorCode::kInvalidArgument,
                      std::string(what) + " failed with GPB error " + std::to_string(rc));
}
static int iso_of(const F64& ell) { return ell.element_count() == 1 ? 1 : 0; }
// operand -> result for the in-place entry points: a no-op under input_output_aliases, a device copy otherwise
static ffi::Error adopt(cudaStream_t stream, const F64& in, ffi::Result<F64>& out) {
    if (out->typed_data() == in.typed_data()) return ffi::Error::Success();
    if (out->size_bytes() != in.size_bytes()) return ffi::Error::InvalidArgument("aliased operand / result sizes differ");
    return cudaMemcpyAsync(out->typed_data(), in.typed_data(), in.size_bytes(), cudaMemcpyDeviceToDevice, stream) == cudaSuccess
               ? ffi::Error::Success()
               : ffi::Error::Internal("cudaMemcpyAsync failed");
}
#define GPB_RETURN_IF_ERROR(expr)          \
    do {                                   \
        ffi::Error e__ = (expr);           \
        if (!e__.success()) return e__;    \
    } while (0)

// ---- K1: K = gram(X, Z)   (kernel.gram / kernel.cross_covariance, gpjax/kernels/computations/dense.py:32-36) -------------
static ffi::Error GramImpl(cudaStream_t stream, F64 X, F64 Z, F64 ell, F64 var, ffi::Result<F64> K, int32_t kind, double diag_add) {
    const int64_t N = X.dimensions()[0], D = X.dimensions()[1], M = Z.dimensions()[0];
    return to_error(gpb_gram(stream, kind, N, M, (int)D, X.typed_data(), D, Z.typed_data(), D, ell.typed_data(), iso_of(ell),
                             var.typed_data(), diag_add, nullptr, 0, K->typed_data(), M),
                    "gpb_gram");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbGram, GramImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<double>("diag_add"));

// VJP of K1 (what jax.grad derives through the double vmap, dense.py:35): cotangents of lengthscale, variance, X, Z
static ffi::Error GramBwdImpl(cudaStream_t stream, F64 X, F64 Z, F64 ell, F64 var, F64 dK, ffi::Result<F64> g_ell,
                              ffi::Result<F64> g_var, ffi::Result<F64> g_X, ffi::Result<F64> g_Z, ffi::Result<F64> ws, int32_t kind) {
    const int64_t N = X.dimensions()[0], D = X.dimensions()[1], M = Z.dimensions()[0];
    // gpb_gram_bwd accumulates: zero the cotangents first
    cudaMemsetAsync(g_ell->typed_data(), 0, g_ell->size_bytes(), stream);
    cudaMemsetAsync(g_var->typed_data(), 0, g_var->size_bytes(), stream);
    cudaMemsetAsync(g_X->typed_data(), 0, g_X->size_bytes(), stream);
    cudaMemsetAsync(g_Z->typed_data(), 0, g_Z->size_bytes(), stream);
    return to_error(gpb_gram_bwd(stream, kind, N, M, (int)D, X.typed_data(), D, Z.typed_data(), D, ell.typed_data(), iso_of(ell),
                                 var.typed_data(), dK.typed_data(), M, 1.0, ws->typed_data(), (int64_t)ws->size_bytes(),
                                 g_ell->typed_data(), g_var->typed_data(), g_X->typed_data(), D, g_Z->typed_data(), D),
                    "gpb_gram_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbGramBwd, GramBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Attr<int32_t>("kind"));

// ---- K2: L = lower_cholesky(A)   (jnp.linalg.cholesky, gpjax/linalg/operations.py:54-55) ------------------------------
// results: L (aliases operand 0), the factorisation workspace (diagonal-block inverses: needed by the solves), info
static ffi::Error PotrfImpl(cudaStream_t stream, F64 A, ffi::Result<F64> L, ffi::Result<F64> ws, ffi::Result<S32> info,
                            int32_t with_potri) {
    const int64_t N = A.dimensions()[0];
    GPB_RETURN_IF_ERROR(adopt(stream, A, L));
    cudaMemsetAsync(info->typed_data(), 0, sizeof(int32_t), stream);
    return to_error(gpb_potrf_lower(stream, N, L->typed_data(), N, 3 /* zero upper + symmetrize_input */, ws->typed_data(), (int64_t)ws->size_bytes(), N, 1, with_potri,
                                    info->typed_data()),
                    "gpb_potrf_lower");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbPotrfLower, PotrfImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<S32>().Attr<int32_t>("with_potri"));

// ---- K3: x = solve(Triangular(L) or its transpose, b)   (jsp.linalg.solve_triangular, operations.py:105-107) ----------
// b is [N] or [N, T]; x aliases operand 1, and the workspace (the solves keep a scratch vector / panel in it next to the
// diagonal-block inverses they read) aliases operand 2
static ffi::Error TrsImpl(cudaStream_t stream, F64 L, F64 b, F64 ws, ffi::Result<F64> x, ffi::Result<F64> ws_out, int32_t trans,
                          int32_t with_potri) {
    const int64_t N = L.dimensions()[0];
    GPB_RETURN_IF_ERROR(adopt(stream, b, x));
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    if (b.dimensions().size() == 1)
        return to_error(gpb_trsv_lower(stream, N, L.typed_data(), N, trans, x->typed_data(), ws_out->typed_data(),
                                       (int64_t)ws_out->size_bytes(), N, 1, with_potri),
                        "gpb_trsv_lower");
    const int64_t T = b.dimensions()[1];
    return to_error(gpb_trsm_lower_left(stream, N, T, L.typed_data(), N, trans, x->typed_data(), T, ws_out->typed_data(),
                                        (int64_t)ws_out->size_bytes(), N, 1, with_potri),
                    "gpb_trsm_lower_left");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbTriangularSolve, TrsImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("trans").Attr<int32_t>("with_potri"));

// ---- K4: logdet(Triangular(L)) = sum(log(diag))   (operations.py:142-144) -----------------------------------------------
static ffi::Error SumLogDiagImpl(cudaStream_t stream, F64 L, ffi::Result<F64> out) {
    const int64_t N = L.dimensions()[0];
    return to_error(gpb_sum_log_diag(stream, N, L.typed_data(), N, out->typed_data()), "gpb_sum_log_diag");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbSumLogDiag, SumLogDiagImpl, ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Ret<F64>());

// ---- K5: Sigma^-1 from the factor (what reverse mode of slogdet / solve materialises) ------------------------------------
// Lbuf (aliases operand 0) is used as scratch above its block diagonal, ws (aliases operand 1) receives the diagonal blocks
static ffi::Error PotriImpl(cudaStream_t stream, F64 L, F64 ws, ffi::Result<F64> Lbuf, ffi::Result<F64> ws_out, ffi::Result<F64> Sinv) {
    const int64_t N = L.dimensions()[0];
    GPB_RETURN_IF_ERROR(adopt(stream, L, Lbuf));
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    return to_error(gpb_potri_lower(stream, N, Lbuf->typed_data(), N, Sinv->typed_data(), N, ws_out->typed_data(),
                                    (int64_t)ws_out->size_bytes(), N, 1, 1),
                    "gpb_potri_lower");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbPotriLower, PotriImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());

// ---- conjugate_mll (gpjax/objectives.py:93-107): forward; value, alpha and the residuals (Sigma/L buffer, workspace) ----------
static ffi::Error MllFwdImpl(cudaStream_t stream, F64 X, F64 y, F64 ell, F64 var, F64 sn, F64 mean, ffi::Result<F64> value,
                             ffi::Result<F64> alpha, ffi::Result<F64> sigma, ffi::Result<F64> ws, ffi::Result<S32> info,
                             int32_t kind, double jitter) {
    const int64_t N = X.dimensions()[0], D = X.dimensions()[1];
    cudaMemsetAsync(info->typed_data(), 0, sizeof(int32_t), stream);
    return to_error(gpb_mll_forward(stream, kind, N, (int)D, X.typed_data(), D, y.typed_data(), ell.typed_data(), iso_of(ell),
                                    var.typed_data(), sn.typed_data(), mean.typed_data(), jitter, sigma->typed_data(), N,
                                    ws->typed_data(), (int64_t)ws->size_bytes(), value->typed_data(), alpha->typed_data(),
                                    info->typed_data()),
                    "gpb_mll_forward");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbMllFwd, MllFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<S32>()
                                  .Attr<int32_t>("kind").Attr<double>("jitter"));

// conjugate_mll VJP: TRTRI + LAUUM run IN PLACE in the residuals, so sigma / ws are operands 4 / 5 AND results 4 / 5
// (input_output_aliases = {4: 4, 5: 5} on the Python side; copied here if XLA did not donate them)
static ffi::Error MllBwdImpl(cudaStream_t stream, F64 X, F64 ell, F64 var, F64 sn, F64 sigma, F64 ws, F64 alpha, F64 gout,
                             ffi::Result<F64> g_ell, ffi::Result<F64> g_var, ffi::Result<F64> g_sn, ffi::Result<F64> g_mean,
                             ffi::Result<F64> sigma_out, ffi::Result<F64> ws_out, int32_t kind) {
    const int64_t N = X.dimensions()[0], D = X.dimensions()[1];
    GPB_RETURN_IF_ERROR(adopt(stream, sigma, sigma_out));
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    return to_error(gpb_mll_backward(stream, kind, N, (int)D, X.typed_data(), D, ell.typed_data(), iso_of(ell), var.typed_data(),
                                     sn.typed_data(), sigma_out->typed_data(), N, ws_out->typed_data(), (int64_t)ws_out->size_bytes(),
                                     alpha.typed_data(), gout.typed_data(), g_ell->typed_data(), g_var->typed_data(),
                                     g_sn->typed_data(), g_mean->typed_data()),
                    "gpb_mll_backward");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbMllBwd, MllBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("kind"));

// ---- collapsed_elbo (gpjax/objectives.py:321-416), the six protocol steps; steps 2 and 5 are jax.lax.psum in Python ----------
// The workspace threads through the steps as operand -> aliased result (every step reads what the previous one left in it).
static ffi::Error SgprStatsImpl(cudaStream_t stream, F64 X, F64 y, F64 Z, F64 ell, F64 var, F64 sn, F64 mean, F64 ws,
                                ffi::Result<F64> Paug, ffi::Result<F64> ws_out, int32_t kind, double jitter, int64_t block_rows,
                                int32_t raw) {
    const int64_t Nloc = X.dimensions()[0], D = X.dimensions()[1], M = Z.dimensions()[0];
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    auto fn = raw ? gpb_sgpr_stats_raw : gpb_sgpr_stats;
    return to_error(fn(stream, kind, Nloc, M, (int)D, X.typed_data(), D, y.typed_data(), Z.typed_data(), D, ell.typed_data(), iso_of(ell),
                       var.typed_data(), sn.typed_data(), mean.typed_data(), jitter, block_rows, ws_out->typed_data(),
                       (int64_t)ws_out->size_bytes(), Paug->typed_data()),
                    raw ? "gpb_sgpr_stats_raw" : "gpb_sgpr_stats");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbSgprStats, SgprStatsImpl,
                              ffi::Ffi::Bind().Ctx<Stream>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<double>("jitter").Attr<int64_t>("block_rows").Attr<int32_t>("raw"));

// need_grad is the flag word of include/gpjax_b200.h: bit 0 = prepare the gradient pass, bit 1 = GPB_FINISH_DENSE_INT8
static ffi::Error SgprFinishImpl(cudaStream_t stream, F64 Z, F64 ell, F64 var, F64 sn, F64 Paug, F64 ws, ffi::Result<F64> elbo,
                                 ffi::Result<S32> info, ffi::Result<F64> ws_out, int32_t kind, int64_t block_rows, int32_t need_grad) {
    const int64_t M = Z.dimensions()[0], D = Z.dimensions()[1];
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    cudaMemsetAsync(info->typed_data(), 0, 2 * sizeof(int32_t), stream);
    return to_error(gpb_sgpr_finish(stream, kind, M, (int)D, Z.typed_data(), D, ell.typed_data(), iso_of(ell), var.typed_data(),
                                    sn.typed_data(), block_rows, ws_out->typed_data(), (int64_t)ws_out->size_bytes(), Paug.typed_data(),
                                    need_grad, elbo->typed_data(), info->typed_data()),
                    "gpb_sgpr_finish");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbSgprFinish, SgprFinishImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<S32>().Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<int64_t>("block_rows").Attr<int32_t>("need_grad"));

static ffi::Error SgprGradLocalImpl(cudaStream_t stream, F64 X, F64 y, F64 Z, F64 ell, F64 var, F64 sn, F64 mean, F64 ws,
                                    ffi::Result<F64> g_Z, ffi::Result<F64> g_ell, ffi::Result<F64> g_var, ffi::Result<F64> ws_out,
                                    int32_t kind, int64_t block_rows) {
    const int64_t Nloc = X.dimensions()[0], D = X.dimensions()[1], M = Z.dimensions()[0];
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    return to_error(gpb_sgpr_grad_local(stream, kind, Nloc, M, (int)D, X.typed_data(), D, y.typed_data(), Z.typed_data(), D,
                                        ell.typed_data(), iso_of(ell), var.typed_data(), sn.typed_data(), mean.typed_data(), block_rows,
                                        ws_out->typed_data(), (int64_t)ws_out->size_bytes(), g_Z->typed_data(), g_ell->typed_data(),
                                        g_var->typed_data()),
                    "gpb_sgpr_grad_local");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbSgprGradLocal, SgprGradLocalImpl,
                              ffi::Ffi::Bind().Ctx<Stream>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<int64_t>("block_rows"));

// step 6: the (all-reduced) partial cotangents come in as operands 6..8 and leave completed as results 0..2 (aliased); the
// workspace (adjoints left by the forward finish + scratch) is operand 4 -> result 5
static ffi::Error SgprGradFinishImpl(cudaStream_t stream, F64 Z, F64 ell, F64 var, F64 sn, F64 ws, F64 gout, F64 g_Z_in,
                                     F64 g_ell_in, F64 g_var_in, ffi::Result<F64> g_Z, ffi::Result<F64> g_ell, ffi::Result<F64> g_var,
                                     ffi::Result<F64> g_sn, ffi::Result<F64> g_mean, ffi::Result<F64> ws_out, int32_t kind,
                                     int64_t block_rows) {
    const int64_t M = Z.dimensions()[0], D = Z.dimensions()[1];
    GPB_RETURN_IF_ERROR(adopt(stream, g_Z_in, g_Z));
    GPB_RETURN_IF_ERROR(adopt(stream, g_ell_in, g_ell));
    GPB_RETURN_IF_ERROR(adopt(stream, g_var_in, g_var));
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    return to_error(gpb_sgpr_grad_finish(stream, kind, M, (int)D, Z.typed_data(), D, ell.typed_data(), iso_of(ell), var.typed_data(),
                                         sn.typed_data(), block_rows, ws_out->typed_data(), (int64_t)ws_out->size_bytes(),
                                         gout.typed_data(), g_Z->typed_data(), g_ell->typed_data(), g_var->typed_data(),
                                         g_sn->typed_data(), g_mean->typed_data()),
                    "gpb_sgpr_grad_finish");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbSgprGradFinish, SgprGradFinishImpl,
                              ffi::Ffi::Bind().Ctx<Stream>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<int64_t>("block_rows"));

// ---- SVGP minibatch elbo (gpjax/objectives.py:241-315): the replicated M x M steps; the streamed steps are the SGPR handlers ----
static ffi::Error SvgpFinishImpl(cudaStream_t stream, F64 Z, F64 ell, F64 var, F64 sn, F64 mean, F64 mu, F64 W, F64 Paug, F64 ws,
                                 ffi::Result<F64> elbo, ffi::Result<S32> info, ffi::Result<F64> ws_out, int32_t kind,
                                 double num_datapoints, double jitter, int64_t block_rows, int32_t need_grad) {
    const int64_t M = Z.dimensions()[0], D = Z.dimensions()[1];
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    cudaMemsetAsync(info->typed_data(), 0, 2 * sizeof(int32_t), stream);
    return to_error(gpb_svgp_finish(stream, kind, M, (int)D, Z.typed_data(), D, ell.typed_data(), iso_of(ell), var.typed_data(),
                                    sn.typed_data(), mean.typed_data(), mu.typed_data(), W.typed_data(), M, num_datapoints, jitter,
                                    block_rows, ws_out->typed_data(), (int64_t)ws_out->size_bytes(), Paug.typed_data(), need_grad,
                                    elbo->typed_data(), info->typed_data()),
                    "gpb_svgp_finish");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbSvgpFinish, SvgpFinishImpl,
                              ffi::Ffi::Bind().Ctx<Stream>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<S32>().Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<double>("num_datapoints").Attr<double>("jitter")
                                  .Attr<int64_t>("block_rows").Attr<int32_t>("need_grad"));

static ffi::Error SvgpGradFinishImpl(cudaStream_t stream, F64 Z, F64 ell, F64 var, F64 sn, F64 W, F64 ws, F64 gout, F64 g_Z_in,
                                     F64 g_ell_in, F64 g_var_in, ffi::Result<F64> g_Z, ffi::Result<F64> g_ell, ffi::Result<F64> g_var,
                                     ffi::Result<F64> g_sn, ffi::Result<F64> g_mean, ffi::Result<F64> g_mu, ffi::Result<F64> g_W,
                                     ffi::Result<F64> ws_out, int32_t kind, double jitter, int64_t block_rows) {
    const int64_t M = Z.dimensions()[0], D = Z.dimensions()[1];
    GPB_RETURN_IF_ERROR(adopt(stream, g_Z_in, g_Z));
    GPB_RETURN_IF_ERROR(adopt(stream, g_ell_in, g_ell));
    GPB_RETURN_IF_ERROR(adopt(stream, g_var_in, g_var));
    GPB_RETURN_IF_ERROR(adopt(stream, ws, ws_out));
    return to_error(gpb_svgp_grad_finish(stream, kind, M, (int)D, Z.typed_data(), D, ell.typed_data(), iso_of(ell), var.typed_data(),
                                         sn.typed_data(), jitter, block_rows, ws_out->typed_data(), (int64_t)ws_out->size_bytes(), gout.typed_data(),
                                         W.typed_data(), M, g_Z->typed_data(), g_ell->typed_data(), g_var->typed_data(),
                                         g_sn->typed_data(), g_mean->typed_data(), g_mu->typed_data(), g_W->typed_data(), M),
                    "gpb_svgp_grad_finish");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(GpbSvgpGradFinish, SvgpGradFinishImpl,
                              ffi::Ffi::Bind().Ctx<Stream>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("kind").Attr<double>("jitter").Attr<int64_t>("block_rows"));
#endif  // GPB_HAVE_XLA_FFI
