// extern "C" boundary (see include/gpjax_b200.h for the contract and the reference citations).
#include "../../include/gpjax_b200.h"
#include "algorithms.h"
#include "sgpr.h"

using namespace gpb;

namespace {
inline int carve(void* ws, int64_t bytes, int64_t n, int d, int potri, FactorWs* out) {
    return factor_ws_carve(ws, bytes, n, d, potri, out);
}
}  // namespace

extern "C" {

const char* gpb_version(void) { return "gpjax_b200 0.1.0 (sm_100a, fp64 DMMA)"; }
int gpb_max_input_dim(void) { return max_input_dim(); }
int64_t gpb_block_size(void) { return block_size_for(0); }
int64_t gpb_block_size_for(int64_t ws_n) { return block_size_for(ws_n); }
void gpb_profile_reset(int enable) { profile_reset(enable); }
void gpb_debug_set_gemm_variant(int v) { debug_set_gemm_variant(v); }
int gpb_profile_read(double* gemm_ms, int64_t* gemm_launches, int64_t* all_launches) {
    return profile_read(gemm_ms, gemm_launches, all_launches);
}

int gpb_gram(void* stream, int kind, int64_t N, int64_t M, int D, const double* X, int64_t ldx, const double* Z,
             int64_t ldz, const double* lengthscale, int lengthscale_is_scalar, const double* variance,
             double diag_add, const double* diag_add_sq, int lower_only, double* K, int64_t ldk) {
    GramDesc g;
    g.kind = kind; g.N = N; g.M = M; g.D = D;
    g.X = X; g.ldx = ldx; g.Z = Z; g.ldz = ldz;
    g.ell = lengthscale; g.ell_is_scalar = lengthscale_is_scalar; g.variance = variance;
    g.K = K; g.ldk = ldk; g.lower_only = lower_only;
    g.diag_add = diag_add; g.diag_add_sq = diag_add_sq;
    return gram(stream, g);
}

int64_t gpb_gram_bwd_workspace_bytes(int64_t N, int64_t M, int D) {
    if (N < 0 || M < 0 || D <= 0) return 0;
    return gram_bwd_partials_count(N, M, D) * (int64_t)sizeof(double);
}

int gpb_gram_bwd(void* stream, int kind, int64_t N, int64_t M, int D, const double* X, int64_t ldx,
                 const double* Z, int64_t ldz, const double* lengthscale, int lengthscale_is_scalar,
                 const double* variance, const double* dK, int64_t lddk, double scale, void* ws, int64_t ws_bytes,
                 double* g_lengthscale, double* g_variance, double* g_X, int64_t ldgx, double* g_Z, int64_t ldgz) {
    if (ws_bytes < gpb_gram_bwd_workspace_bytes(N, M, D)) return GPB_ERR_WORKSPACE;
    GramBwdDesc d;
    d.kind = kind; d.N = N; d.M = M; d.D = D;
    d.X = X; d.ldx = ldx; d.Z = Z; d.ldz = ldz;
    d.ell = lengthscale; d.ell_is_scalar = lengthscale_is_scalar; d.variance = variance;
    d.dK = dK; d.lddk = lddk; d.scale = scale; d.partials = static_cast<double*>(ws);
    d.g_ell = g_lengthscale; d.g_var = g_variance; d.g_X = g_X; d.ldgx = ldgx; d.g_Z = g_Z; d.ldgz = ldgz;
    return gram_bwd(stream, d);
}

int64_t gpb_factor_workspace_bytes(int64_t ws_n, int ws_d, int ws_potri) {
    return factor_ws_bytes(ws_n, ws_d, ws_potri);
}

int gpb_potrf_lower(void* stream, int64_t N, double* A, int64_t lda, int zero_upper, void* ws, int64_t ws_bytes,
                    int64_t ws_n, int ws_d, int ws_potri, int* info) {
    if (N > ws_n) return GPB_ERR_WORKSPACE;
    FactorWs w;
    int rc = carve(ws, ws_bytes, ws_n, ws_d, ws_potri, &w);
    if (rc) return rc;
    if ((rc = factor_set_planes(stream, w, N, nullptr, nullptr, 0.0))) return rc;  // bare matrix: 7 planes unless forced
    if ((zero_upper & 2) && (rc = symmetrize_average_lower(stream, N, A, lda))) return rc;
    rc = potrf_lower(stream, N, A, lda, w, info);
    if (rc) return rc;
    if (zero_upper & 1) return zero_triangle(stream, N, A, lda, 2);
    return GPB_OK;
}

int gpb_diag_inverses(void* stream, int64_t N, const double* L, int64_t lda, void* ws, int64_t ws_bytes,
                      int64_t ws_n, int ws_d, int ws_potri) {
    if (N > ws_n) return GPB_ERR_WORKSPACE;
    FactorWs w;
    int rc = carve(ws, ws_bytes, ws_n, ws_d, ws_potri, &w);
    if (rc) return rc;
    return diag_inverses(stream, N, L, lda, w);
}

int gpb_trsv_lower(void* stream, int64_t N, const double* L, int64_t lda, int trans, double* x, void* ws,
                   int64_t ws_bytes, int64_t ws_n, int ws_d, int ws_potri) {
    if (N > ws_n) return GPB_ERR_WORKSPACE;
    FactorWs w;
    int rc = carve(ws, ws_bytes, ws_n, ws_d, ws_potri, &w);
    if (rc) return rc;
    return trsv_lower(stream, N, L, lda, w, x, trans);
}

int gpb_trsm_lower_left(void* stream, int64_t N, int64_t T, const double* L, int64_t lda, int trans, double* B,
                        int64_t ldb, void* ws, int64_t ws_bytes, int64_t ws_n, int ws_d, int ws_potri) {
    if (N > ws_n || T > ws_n) return GPB_ERR_WORKSPACE;
    FactorWs w;
    int rc = carve(ws, ws_bytes, ws_n, ws_d, ws_potri, &w);
    if (rc) return rc;
    return trsm_lower_left(stream, N, T, L, lda, w, B, ldb, trans);
}

int gpb_sum_log_diag(void* stream, int64_t N, const double* L, int64_t lda, double* out) {
    if (N < 0 || !out || (N > 0 && !L)) return GPB_ERR_INVALID;
    return sum_log_diag(stream, N, L, lda, out);
}

int gpb_potri_lower(void* stream, int64_t N, double* A, int64_t lda, double* out, int64_t ldo, void* ws,
                    int64_t ws_bytes, int64_t ws_n, int ws_d, int ws_potri) {
    if (N > ws_n || !ws_potri) return GPB_ERR_WORKSPACE;
    FactorWs w;
    int rc = carve(ws, ws_bytes, ws_n, ws_d, ws_potri, &w);
    if (rc) return rc;
    if ((rc = factor_set_planes(stream, w, N, nullptr, nullptr, 0.0))) return rc;
    if ((rc = trtri_into_upper(stream, N, A, lda, w))) return rc;
    if ((rc = lauum_upper(stream, N, A, lda, w))) return rc;
    const int64_t NB = w.nb, nblk = nblocks(N, NB);
    for (int64_t k = 0; k < nblk; ++k) {
        const int64_t j0 = k * NB;
        const int64_t nbk = (N - j0) < NB ? (N - j0) : NB;
        if ((rc = copy2d(stream, nbk, nbk, w.Sdiag + k * NB * NB, NB, out + j0 * ldo + j0, ldo))) return rc;
        const int64_t right = N - j0 - nbk;
        if (right > 0 && (rc = copy2d(stream, nbk, right, A + j0 * lda + j0 + nbk, lda, out + j0 * ldo + j0 + nbk, ldo)))
            return rc;
    }
    return symmetrize(stream, N, out, ldo, 0);
}

int gpb_gemm(void* stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
             int a_layout, const double* B, int64_t ldb, int b_layout, double beta, double* C, int64_t ldc,
             int mask) {
    GemmDesc g;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = lda; g.a_layout = a_layout;
    g.B = B; g.ldb = ldb; g.b_layout = b_layout;
    g.C = C; g.ldc = ldc; g.alpha = alpha; g.beta = beta; g.mask = mask;
    return gemm(stream, g);
}

int gpb_profile_read_ozaki(double* ms, int64_t* launches, double* int8_ops) { return profile_read_ozaki(ms, launches, int8_ops); }
int gpb_ozaki_available(void) { return ozaki_available() ? 1 : 0; }
void gpb_set_ozaki_slices(int nslices) { set_ozaki_slices(nslices); }
int gpb_get_ozaki_slices(void) { return get_ozaki_slices(); }
int gpb_ozaki_auto_planes(int64_t N, double variance, double obs_stddev, double jitter) {
    return ozaki_auto_planes_host(N, variance, obs_stddev, jitter);
}
int gpb_ozaki_slice(void* stream, int64_t rows, int64_t K, const double* X, int64_t ldx, int nslices, void* Q,
                    int64_t ldq, double* scale) {
    return ozaki_slice(stream, rows, K, K, X, ldx, nslices, static_cast<int8_t*>(Q), ldq, scale);
}
int gpb_ozaki_gemm(void* stream, int64_t M, int64_t N, int64_t K, int nslices, const void* Qa, int64_t ldqa,
                   const double* scale_a, const void* Qb, int64_t ldqb, const double* scale_b, double alpha,
                   double* C, int64_t ldc, int mask_lower) {
    OzakiGemmDesc d;
    d.M = M; d.N = N; d.K = K; d.nslices = nslices;
    d.Qa = static_cast<const int8_t*>(Qa); d.ldqa = ldqa; d.sa = scale_a;
    d.Qb = static_cast<const int8_t*>(Qb); d.ldqb = ldqb; d.sb = scale_b;
    d.alpha = alpha; d.C = C; d.ldc = ldc; d.mask = mask_lower ? MASK_LOWER : MASK_NONE;
    return ozaki_gemm(stream, d);
}
int gpb_igemm_i8(void* stream, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                 int64_t ldb, void* C, int64_t ldc) {
    return igemm_i8(stream, M, N, K, static_cast<const int8_t*>(A), lda, static_cast<const int8_t*>(B), ldb,
                    static_cast<int32_t*>(C), ldc);
}

int64_t gpb_mll_workspace_bytes(int64_t N, int D) { return factor_ws_bytes(N, D, 1); }

int gpb_mll_forward(void* stream, int kind, int64_t N, int D, const double* X, int64_t ldx, const double* y,
                    const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                    const double* obs_stddev, const double* mean_const, double jitter, double* Sigma, int64_t lds,
                    void* ws, int64_t ws_bytes, double* value_out, double* alpha_out, int* info) {
    FactorWs w;
    int rc = carve(ws, ws_bytes, N, D, 1, &w);
    if (rc) return rc;
    MllArgs a;
    a.kind = kind; a.N = N; a.D = D; a.X = X; a.ldx = ldx; a.y = y;
    a.ell = lengthscale; a.ell_is_scalar = lengthscale_is_scalar; a.variance = variance;
    a.obs_stddev = obs_stddev; a.mean_const = mean_const; a.jitter = jitter; a.Sigma = Sigma; a.lds = lds;
    return mll_forward(stream, a, w, value_out, alpha_out, info);
}

int gpb_mll_backward(void* stream, int kind, int64_t N, int D, const double* X, int64_t ldx,
                     const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                     const double* obs_stddev, double* Sigma, int64_t lds, void* ws, int64_t ws_bytes,
                     const double* alpha, const double* gout, double* g_lengthscale, double* g_variance,
                     double* g_obs_stddev, double* g_mean_const) {
    FactorWs w;
    int rc = carve(ws, ws_bytes, N, D, 1, &w);
    if (rc) return rc;
    MllArgs a;
    a.kind = kind; a.N = N; a.D = D; a.X = X; a.ldx = ldx;
    a.ell = lengthscale; a.ell_is_scalar = lengthscale_is_scalar; a.variance = variance;
    a.obs_stddev = obs_stddev; a.Sigma = Sigma; a.lds = lds;
    return mll_backward(stream, a, w, alpha, gout, g_lengthscale, g_variance, g_obs_stddev, g_mean_const);
}

int64_t gpb_sgpr_workspace_bytes(int64_t M, int D, int64_t block_rows) { return sgpr_ws_bytes(M, D, block_rows); }
int64_t gpb_sgpr_stats_count(int64_t M) { return (M + 2) * (M + 2); }

static SgprArgs sgpr_args(int kind, int64_t Nloc, int64_t M, int D, const double* X, int64_t ldx, const double* y,
                          const double* Z, int64_t ldz, const double* ell, int iso, const double* var,
                          const double* sn, const double* mean, double jitter, int64_t block_rows) {
    SgprArgs a;
    a.kind = kind; a.Nloc = Nloc; a.M = M; a.D = D; a.X = X; a.ldx = ldx; a.y = y; a.Z = Z; a.ldz = ldz;
    a.ell = ell; a.ell_is_scalar = iso; a.variance = var; a.obs_stddev = sn; a.mean_const = mean;
    a.jitter = jitter; a.block_rows = block_rows;
    return a;
}

int gpb_sgpr_stats(void* stream, int kind, int64_t Nloc, int64_t M, int D, const double* X, int64_t ldx,
                   const double* y, const double* Z, int64_t ldz, const double* lengthscale,
                   int lengthscale_is_scalar, const double* variance, const double* obs_stddev,
                   const double* mean_const, double jitter, int64_t block_rows, void* ws, int64_t ws_bytes,
                   double* Paug) {
    SgprWs w;
    int rc = sgpr_ws_carve(ws, ws_bytes, M, D, block_rows, &w);
    if (rc) return rc;
    return sgpr_stats(stream, sgpr_args(kind, Nloc, M, D, X, ldx, y, Z, ldz, lengthscale, lengthscale_is_scalar,
                                        variance, obs_stddev, mean_const, jitter, block_rows), w, Paug);
}

int gpb_sgpr_stats_raw(void* stream, int kind, int64_t Nloc, int64_t M, int D, const double* X, int64_t ldx,
                       const double* y, const double* Z, int64_t ldz, const double* lengthscale,
                       int lengthscale_is_scalar, const double* variance, const double* obs_stddev,
                       const double* mean_const, double jitter, int64_t block_rows, void* ws, int64_t ws_bytes,
                       double* Paug) {
    SgprWs w;
    int rc = sgpr_ws_carve(ws, ws_bytes, M, D, block_rows, &w);
    if (rc) return rc;
    SgprArgs a = sgpr_args(kind, Nloc, M, D, X, ldx, y, Z, ldz, lengthscale, lengthscale_is_scalar, variance, obs_stddev,
                           mean_const, jitter, block_rows);
    a.raw_stats = 1;
    a.dense_int8 = 1;  // the whitening of the raw sums: same precondition as the route itself
    return sgpr_stats(stream, a, w, Paug);
}

int gpb_sgpr_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz,
                    const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                    const double* obs_stddev, int64_t block_rows, void* ws, int64_t ws_bytes, const double* Paug,
                    int need_grad, double* elbo_out, int* info_out) {
    SgprWs w;
    int rc = sgpr_ws_carve(ws, ws_bytes, M, D, block_rows, &w);
    if (rc) return rc;
    SgprArgs a = sgpr_args(kind, 0, M, D, nullptr, 0, nullptr, Z, ldz, lengthscale, lengthscale_is_scalar, variance, obs_stddev,
                           nullptr, 0.0, block_rows);
    a.dense_int8 = (need_grad & GPB_FINISH_DENSE_INT8) ? 1 : 0;
    return sgpr_finish(stream, a, w, Paug, need_grad & 1, elbo_out, info_out);
}

int gpb_sgpr_grad_local(void* stream, int kind, int64_t Nloc, int64_t M, int D, const double* X, int64_t ldx,
                        const double* y, const double* Z, int64_t ldz, const double* lengthscale,
                        int lengthscale_is_scalar, const double* variance, const double* obs_stddev,
                        const double* mean_const, int64_t block_rows, void* ws, int64_t ws_bytes, double* g_Z,
                        double* g_lengthscale, double* g_variance) {
    SgprWs w;
    int rc = sgpr_ws_carve(ws, ws_bytes, M, D, block_rows, &w);
    if (rc) return rc;
    return sgpr_grad_local(stream, sgpr_args(kind, Nloc, M, D, X, ldx, y, Z, ldz, lengthscale, lengthscale_is_scalar,
                                             variance, obs_stddev, mean_const, 0.0, block_rows), w, g_Z,
                           g_lengthscale, g_variance);
}

int gpb_sgpr_grad_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz,
                         const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                         const double* obs_stddev, int64_t block_rows, void* ws, int64_t ws_bytes, const double* gout,
                         double* g_Z, double* g_lengthscale, double* g_variance, double* g_obs_stddev,
                         double* g_mean_const) {
    SgprWs w;
    int rc = sgpr_ws_carve(ws, ws_bytes, M, D, block_rows, &w);
    if (rc) return rc;
    return sgpr_grad_finish(stream, sgpr_args(kind, 0, M, D, nullptr, 0, nullptr, Z, ldz, lengthscale,
                                              lengthscale_is_scalar, variance, obs_stddev, nullptr, 0.0, block_rows),
                            w, gout, g_Z, g_lengthscale, g_variance, g_obs_stddev, g_mean_const);
}

int gpb_svgp_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz, const double* lengthscale,
                    int lengthscale_is_scalar, const double* variance, const double* obs_stddev,
                    const double* mean_const, const double* mu, const double* W, int64_t ldw, double num_datapoints,
                    double jitter, int64_t block_rows, void* ws, int64_t ws_bytes, const double* Paug, int need_grad,
                    double* elbo_out, int* info_out) {
    SgprWs w;
    int rc = sgpr_ws_carve(ws, ws_bytes, M, D, block_rows, &w);
    if (rc) return rc;
    SgprArgs a = sgpr_args(kind, 0, M, D, nullptr, 0, nullptr, Z, ldz, lengthscale, lengthscale_is_scalar, variance, obs_stddev,
                           mean_const, jitter, block_rows);
    a.dense_int8 = (need_grad & GPB_FINISH_DENSE_INT8) ? 1 : 0;
    return svgp_finish(stream, a, w, Paug, mu, W, ldw, num_datapoints, need_grad & 1, elbo_out, info_out);
}

int gpb_svgp_grad_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz,
                         const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                         const double* obs_stddev, double jitter, int64_t block_rows, void* ws, int64_t ws_bytes,
                         const double* gout, const double* W, int64_t ldw, double* g_Z, double* g_lengthscale,
                         double* g_variance, double* g_obs_stddev, double* g_mean_const, double* g_mu, double* g_W,
                         int64_t ldgw) {
    SgprWs w;
    int rc = sgpr_ws_carve(ws, ws_bytes, M, D, block_rows, &w);
    if (rc) return rc;
    return svgp_grad_finish(stream, sgpr_args(kind, 0, M, D, nullptr, 0, nullptr, Z, ldz, lengthscale,
                                              lengthscale_is_scalar, variance, obs_stddev, nullptr, jitter, block_rows),
                            w, gout, W, ldw, g_Z, g_lengthscale, g_variance, g_obs_stddev, g_mean_const, g_mu, g_W, ldgw);
}

}  // extern "C"
