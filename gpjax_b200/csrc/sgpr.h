// collapsed_elbo (SGPR) through row-additive statistics -- gpjax/objectives.py:321-416.
// Host orchestration only (device pointers are opaque), same rules as algorithms.h.
#pragma once
#include "algorithms.h"

namespace gpb {

struct SgprArgs {
    int kind = 0;
    int64_t Nloc = 0;  // rows of this rank's shard
    int64_t M = 0;     // inducing points
    int D = 0;
    const double* X = nullptr;  // [Nloc, D]
    int64_t ldx = 0;
    const double* y = nullptr;  // [Nloc]
    const double* Z = nullptr;  // [M, D]
    int64_t ldz = 0;
    const double* ell = nullptr;
    int ell_is_scalar = 0;
    const double* variance = nullptr;
    const double* obs_stddev = nullptr;
    const double* mean_const = nullptr;  // nullable
    double jitter = 1e-6;
    int64_t block_rows = 32768;  // rows per streamed block
    int raw_stats = 0;           // pass 1 accumulates raw [K_b^T|d|1] products and whitens once at the end (see sgpr.cpp)
    int dense_int8 = 0;          // the caller vouches for a well-conditioned Kzz (the raw-statistics route's own precondition): the dense
                                 // M x M x M products of the replicated finish may run as 7-plane int8 digit products (mm_gemm)
};

struct SgprWs {
    FactorWs fz, fb;
    double *Lz, *Linv, *Bmat, *LB, *Binv, *G1, *G2, *Tmp, *Caug, *dKzz, *Wc, *H, *X3;
    double *psi, *a1, *w, *v, *u, *cvec, *rowsum, *phiu, *tvec, *sc, *dots;
    double *T1, *T2, *Ppart;
    double* gpart;
    int* info2;  // [2]: info of chol(Kzz), chol(B)
    // int8 digit planes for the pass-2 product (null when the block is too small for the int8 path): T1 rows and Caug rows
    int8_t *oz_qt = nullptr, *oz_qc = nullptr;
    double *oz_st = nullptr, *oz_sc = nullptr;
    int64_t oz_kplane = 0;  // digits per plane = M + 2 rounded up to 128
    // pass 1 (statistics SYRK, contraction over the block rows): column digit planes share oz_qt; plane = block rows rounded to 128
    double *oz_s1 = nullptr, *oz_colmax = nullptr, *oz_ones = nullptr;
    int64_t oz_kplane1 = 0;
    // capacities of the digit buffers (bytes) and of the scale vectors (entries): the dense M x M x M products of the replicated
    // finish borrow them between the two streamed passes (mm_gemm in sgpr.cpp)
    int64_t oz_qt_bytes = 0, oz_qc_bytes = 0, oz_st_len = 0, oz_sc_len = 0;
};

int64_t sgpr_ws_bytes(int64_t M, int D, int64_t block_rows);
int sgpr_ws_carve(void* buf, int64_t bytes, int64_t M, int D, int64_t block_rows, SgprWs* ws);
inline int64_t sgpr_stats_ld(int64_t M) { return M + 2; }

// Pass 1 (per rank): Paug[(M+2)x(M+2)] (row stride M+2, lower triangle) = local sums of
//   [A~ ; d^T ; 1^T] [A~ ; d^T ; 1^T]^T  with  A~ = Lz^-1 Kzx  (UNSCALED by the noise).
// It is exactly the quantity that is all-reduced over ranks (M^2 + 2M + 3 useful doubles).
int sgpr_stats(stream_t s, const SgprArgs& a, const SgprWs& ws, double* Paug);
// Replicated M x M finish on the (all-reduced) statistics: ELBO value; with need_grad also the
// adjoint matrices for pass 2 (kept in ws).
int sgpr_finish(stream_t s, const SgprArgs& a, const SgprWs& ws, const double* Paug, int need_grad, double* elbo_out,
                int* info_out);
// Pass 2 (per rank): local contributions  g_Z[M,D], g_ell[D|1], g_var[1]  (overwritten).
int sgpr_grad_local(stream_t s, const SgprArgs& a, const SgprWs& ws, double* g_Z, double* g_ell, double* g_var);
// Replicated part added to the (all-reduced) local gradients: Kzz term + scalar terms; then
// everything is scaled by *gout (null -> 1).  g_obs / g_mean are overwritten.
int sgpr_grad_finish(stream_t s, const SgprArgs& a, const SgprWs& ws, const double* gout, double* g_Z, double* g_ell,
                     double* g_var, double* g_obs, double* g_mean);

// SVGP (uncollapsed ELBO, gpjax/objectives.py:241-315): same pass 1 / pass 2 / all-reduces as SGPR (sgpr_stats,
// sgpr_grad_local); only the replicated M x M finish differs.  mu: variational mean [M]; W: lower-triangular
// variational root covariance [M x M]; num_datapoints: N of the full data set (the batch size B comes out of Paug).
int svgp_finish(stream_t s, const SgprArgs& a, const SgprWs& ws, const double* Paug, const double* mu, const double* W,
                int64_t ldw, double num_datapoints, int need_grad, double* elbo_out, int* info_out);
int svgp_grad_finish(stream_t s, const SgprArgs& a, const SgprWs& ws, const double* gout, const double* W, int64_t ldw,
                     double* g_Z, double* g_ell, double* g_var, double* g_obs, double* g_mean, double* g_mu, double* g_W,
                     int64_t ldgw);

}  // namespace gpb
