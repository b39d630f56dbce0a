/* gpjax_b200 -- C ABI of the B200-native GPJax hot path (libgpjax_b200.so).
 *
 * Binding-agnostic boundary: every entry point is `extern "C"`, takes plain pointers and sizes,
 * and is what a `jax.ffi` XLA custom-call handler (or the ctypes/torch shim shipped in
 * gpjax_b200/_lib.py) binds for the reference call site cited next to it.  Conventions:
 *
 *   - float64, row-major (C order); `ld*` are row strides in elements.
 *   - every `const double*` / `double*` is a DEVICE pointer owned by the caller, including the
 *     hyper-parameters (lengthscale[D] or [1], variance, obs_stddev, mean constant) so a training
 *     loop never synchronises with the host.
 *   - `stream` is a cudaStream_t.  Calls only enqueue work: no allocation, no synchronisation,
 *     no host read-back, no exceptions.  Re-entrant: one process may drive several devices from several
 *     threads (XLA's per-device executors) -- everything a call mutates lives in the buffers it is handed;
 *     library-internal launch state (kernel attributes, the look-ahead helper streams and their events) is kept
 *     PER DEVICE behind a mutex, and the only process-wide words are the atomic configuration switches
 *     (gpb_set_ozaki_slices, gpb_profile_reset), which are not meant to change while calls are in flight.
 *   - return value: 0 = OK, <0 = GPB_ERR_* (argument / capability / launch error).  Numerical
 *     failure (non-positive-definite pivot) is NOT an error code: outputs are NaN-filled and the
 *     device word `*info` receives the 1-based index of the first failing pivot, mirroring the
 *     NaN semantics of jnp.linalg.cholesky that the reference relies on.
 *   - scratch memory is passed in; sizes come from the companion `*_workspace_bytes` query.
 *     Factorisation workspaces are position-dependent: pass the same (ws, ws_n, ws_d, ws_potri)
 *     to every call that works on the same factor.
 *   - kind: 0 = RBF (gpjax/kernels/stationary/rbf.py:40-44), 1 = Matern32 (matern32.py:41-54),
 *           2 = Matern52 (matern52.py:42-53), 3 = Matern12 (matern12.py:44-48).
 *
 * Citations are file:line in the reference tree (gpjax 0.13.2).
 */
#ifndef GPJAX_B200_H
#define GPJAX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_OK 0
#define GPB_ERR_INVALID (-1)
#define GPB_ERR_UNSUPPORTED (-2)
#define GPB_ERR_LAUNCH (-3)
#define GPB_ERR_WORKSPACE (-4)
#define GPB_FINISH_DENSE_INT8 2 /* flag bit of need_grad in gpb_sgpr_finish / gpb_svgp_finish (see there) */

#define GPB_KIND_RBF 0
#define GPB_KIND_MATERN32 1
#define GPB_KIND_MATERN52 2
#define GPB_KIND_MATERN12 3
/* Kinds with one extra "shape" scalar.  For these every `variance` argument points to TWO consecutive device
 * doubles {variance, shape} and every `g_variance` output to {g_variance, g_shape}. */
#define GPB_KIND_RATIONAL_QUADRATIC 4 /* shape = alpha   (gpjax/kernels/stationary/rational_quadratic.py:77-83) */
#define GPB_KIND_POWERED_EXPONENTIAL 5 /* shape = power  (powered_exponential.py:85-89); not for the sparse objectives */
#define GPB_KIND_PERIODIC 6            /* shape = period  (periodic.py:81-88) */
#define GPB_KIND_WHITE 7               /* variance * all(x == y)  (white.py:63-64); lengthscale must be 1 */

const char* gpb_version(void);
int gpb_max_input_dim(void); /* largest D the compiled kernels accept */
int64_t gpb_block_size(void); /* NB of the blocked algorithms for small orders (= gpb_block_size_for(0)) */
/* NB used by every call that shares a workspace sized for order ws_n (1024 below 12,288 rows, 2048 from there on; measured in
 * gpjax_b200/csrc/algorithms.h).  Reporting only: callers never pass a block size. */
int64_t gpb_block_size_for(int64_t ws_n);

/* Measurement hooks for bench.py (off by default; not part of the reference-facing surface):
 * while enabled, every DMMA GEMM launch is bracketed by CUDA events on its stream and every kernel
 * launch of the library is counted.  gpb_profile_read blocks until the recorded events completed. */
void gpb_profile_reset(int enable);
void gpb_debug_set_gemm_variant(int v); /* kernel-tuning hook for scripts/gemm_bench.py; 0 = default */
int gpb_profile_read(double* gemm_ms, int64_t* gemm_launches, int64_t* all_launches);
/* same for the launches of the int8 (Ozaki) kernel: summed duration, count, algorithmic int8 operations (2 x MACs) */
int gpb_profile_read_ozaki(double* ms, int64_t* launches, double* int8_ops);

/* ---- K1: fused Gram / cross-covariance -------------------------------------------------------
 * Replaces DenseKernelComputation._cross_covariance (gpjax/kernels/computations/dense.py:32-36),
 * AbstractKernelComputation.gram (computations/base.py:56-72) and, through diag_add/diag_add_sq,
 * add_jitter (gpjax/linalg/utils.py:39-65) + "eye * obs_noise" (gpjax/objectives.py:101-102).
 * K[i,j] = variance * g(sum_d (X[i,d]/l_d - Z[j,d]/l_d)^2); where i == j (square use):
 * += diag_add + (*diag_add_sq)^2.  lower_only: skip tiles strictly above the diagonal. */
int gpb_gram(void* stream, int kind, int64_t N, int64_t M, int D, const double* X, int64_t ldx,
             const double* Z, int64_t ldz, const double* lengthscale, int lengthscale_is_scalar,
             const double* variance, double diag_add, const double* diag_add_sq, int lower_only,
             double* K, int64_t ldk);

/* Reverse mode of gpb_gram (what jax.grad derives through the double vmap, dense.py:35):
 * g_lengthscale/g_variance/g_X/g_Z += scale * <dK, dK/d.> ; any g_* may be NULL. */
int64_t gpb_gram_bwd_workspace_bytes(int64_t N, int64_t M, int D);
int gpb_gram_bwd(void* stream, int kind, int64_t N, int64_t M, int D, const double* X, int64_t ldx,
                 const double* Z, int64_t ldz, const double* lengthscale, int lengthscale_is_scalar,
                 const double* variance, const double* dK, int64_t lddk, double scale, void* ws,
                 int64_t ws_bytes, double* g_lengthscale, double* g_variance, double* g_X, int64_t ldgx,
                 double* g_Z, int64_t ldgz);

/* ---- K2..K5: Cholesky / triangular solves / log-determinant -------------------------------------
 * gpb_potrf_lower  : lower_cholesky(Dense) -> jnp.linalg.cholesky (gpjax/linalg/operations.py:54-55);
 *                    in place; `zero_upper` is a flag word: bit 0 zeroes the strict upper triangle as JAX returns
 *                    it, bit 1 first replaces the lower triangle by (A + A^T) / 2 -- jnp.linalg.cholesky's
 *                    symmetrize_input=True default; without bit 1 only the lower triangle is read (the fused
 *                    objectives build Sigma symmetric by construction and never set it).
 * gpb_diag_inverses: prepares the workspace for solves against a factor that was not produced by
 *                    gpb_potrf_lower (a user-built Triangular).
 * gpb_trsv_lower / gpb_trsm_lower_left : solve(Triangular, b) -> jsp.linalg.solve_triangular
 *                    (operations.py:105-107); trans=1 solves with L^T (Triangular.T, operators.py:216-218).
 * gpb_sum_log_diag : logdet(Triangular) = sum(log(diag)) WITHOUT the factor 2 (operations.py:142-144).
 * gpb_potri_lower  : Sigma^-1 from the factor (TRTRI + LAUUM); result mirrored to a full symmetric
 *                    matrix in `out` (ld ldo); what reverse-mode of slogdet/solve materialises. */
int64_t gpb_factor_workspace_bytes(int64_t ws_n, int ws_d, int ws_potri);
int gpb_potrf_lower(void* stream, int64_t N, double* A, int64_t lda, int zero_upper, void* ws,
                    int64_t ws_bytes, int64_t ws_n, int ws_d, int ws_potri, int* info);
int gpb_diag_inverses(void* stream, int64_t N, const double* L, int64_t lda, void* ws, int64_t ws_bytes,
                      int64_t ws_n, int ws_d, int ws_potri);
int gpb_trsv_lower(void* stream, int64_t N, const double* L, int64_t lda, int trans, double* x, void* ws,
                   int64_t ws_bytes, int64_t ws_n, int ws_d, int ws_potri);
int gpb_trsm_lower_left(void* stream, int64_t N, int64_t T, const double* L, int64_t lda, int trans,
                        double* B, int64_t ldb, void* ws, int64_t ws_bytes, int64_t ws_n, int ws_d,
                        int ws_potri);
int gpb_sum_log_diag(void* stream, int64_t N, const double* L, int64_t lda, double* out);
int gpb_potri_lower(void* stream, int64_t N, double* A, int64_t lda, double* out, int64_t ldo, void* ws,
                    int64_t ws_bytes, int64_t ws_n, int ws_d, int ws_potri);

/* C = beta*C + alpha * A * B^T on the FP64 tensor pipe (jnp.matmul call sites, objectives.py:390,404).
 * layouts: 0 = operand stored [rows][K] (K contiguous), 1 = stored [K][rows]. */
int gpb_gemm(void* stream, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
             int a_layout, const double* B, int64_t ldb, int b_layout, double beta, double* C, int64_t ldc,
             int mask);

/* ---- FP64 rank-k updates on the INT8 tensor pipe (Ozaki scheme; tcgen05.mma kind::i8) --------------------------
 * Building blocks of gpb_potrf_lower / gpb_potri_lower, i.e. still jnp.linalg.cholesky (gpjax/linalg/operations.py:54-55)
 * and the inverse its reverse mode needs -- there is no separate reference function.  An fp64 operand row is split
 * into `nslices` (1..7) balanced radix-256 digit planes (every plane uses the whole int8 range, 8 bits per plane) after a
 * power-of-two row scaling (gpb_ozaki_slice); every digit-pair product is an exact int8 x int8 -> int32 GEMM on the tcgen05 pipe
 * and gpb_ozaki_gemm recombines the orders p+q < nslices in fp64:  C += alpha * A B^T  up to a truncation error of
 * 16 (nslices+1) 2^(-8 nslices) K relative to the row maxima (nslices = 7: below fp64 rounding of a K = 1024 product).
 *   Q        : int8 [rows, nslices*K], plane p at columns [p*K, (p+1)*K); 16-byte aligned, ldq multiple of 16
 *   scale    : fp64 [rows], 2^e_i (NaN for a row holding NaN/Inf -> NaN output, JAX semantics)
 *   K        : multiple of 128, and nslices * K * 16384 < 2^31 (int32 headroom of the deepest order; GPB_ERR_UNSUPPORTED beyond --
 *              callers split K, as the SGPR statistics do for K = 65,536)
 * gpb_igemm_i8 exposes the raw integer product (C int32 = A B^T) for bit-exact testing.
 * gpb_set_ozaki_slices(s): s in {-1 (auto, default), 0, 4..7}; 0 keeps every blocked algorithm on the FP64 DMMA pipe, otherwise
 * the rank-NB trailing updates (>= 2048 output rows) of potrf / trtri / lauum and the panel x inverse-diagonal-block products next
 * to them (environment variable GPB_OZ_PANELS=0 keeps those on the DMMA pipe) run through gpb_ozaki_gemm, and the two streamed
 * products of gpb_sgpr_stats(_raw) / gpb_sgpr_grad_local (blocks of >= 2048 rows, M >= 256) with all 7 planes (56 bits).
 * Plane count of the exact-GP updates in auto mode -- the conditioning guard -- is decided per call ON THE DEVICE (a one-thread
 * kernel writes it into the workspace, the product kernels read it: no host synchronisation):
 *   gpb_potrf_lower / gpb_potri_lower (a bare matrix, nothing known about it): 7 planes = fp64-rounding-level products;
 *   gpb_mll_forward / gpb_mll_backward: 6 planes iff the hyper-parameters PROVE cond(Sigma) <= 2e6 through
 *     cond(K + s I) <= (N variance + s) / s, s = obs_stddev^2 + jitter  (|k| <= variance), else 7.
 * Measured against the CPU oracle (profiles/r02_cond_sweep_n8192.jsonl, r02_cond_sweep_radix256.jsonl, r02_cond_sweep_panels_int8.jsonl): 56 bits equal the FP64
 * path's own error at every conditioning; 48 bits carry at most 1.4e-15 x bound relative error in the most sensitive gradient
 * (<= 2.8e-9 under the guard; contract 1e-8).
 * gpb_ozaki_auto_planes is the same rule evaluated on host values, for reporting and tests only.
 * Environment variable GPB_OZAKI ("auto", 0, 4..7) sets the initial value of the switch; the switch is an atomic
 * process-wide configuration word, not meant to change while calls are in flight. */
int gpb_ozaki_available(void);
void gpb_set_ozaki_slices(int nslices);
int gpb_get_ozaki_slices(void);
int gpb_ozaki_auto_planes(int64_t N, double variance, double obs_stddev, double jitter);
int gpb_ozaki_slice(void* stream, int64_t rows, int64_t K, const double* X, int64_t ldx, int nslices, void* Q,
                    int64_t ldq, double* scale);
int gpb_ozaki_gemm(void* stream, int64_t M, int64_t N, int64_t K, int nslices, const void* Qa, int64_t ldqa,
                   const double* scale_a, const void* Qb, int64_t ldqb, const double* scale_b, double alpha,
                   double* C, int64_t ldc, int mask_lower);
int gpb_igemm_i8(void* stream, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                 int64_t ldb, void* C, int64_t ldc);

/* ---- conjugate_mll value + analytic gradient -----------------------------------------------------
 * Forward = gpjax/objectives.py:93-107 + GaussianDistribution.log_prob (gpjax/distributions.py:124-134):
 *   Sigma = K(X,X) + (jitter + obs_stddev^2) I,  value = -1/2 (N log 2pi + logdet Sigma + d^T Sigma^-1 d),
 *   d = y - mean_const.  (The reference evaluates logdet/solve through LU; Cholesky is used here --
 *   identical for SPD Sigma.)  Sigma is an N x N scratch buffer that afterwards holds L (lower).
 * Backward = what jax.value_and_grad (gpjax/fit.py:160) returns for the constrained parameters:
 *   W = 1/2 (alpha alpha^T - Sigma^-1) contracted with dK/dtheta tile by tile.  Must be called with
 *   the Sigma buffer, workspace and alpha exactly as the forward left them.  mean_const may be NULL
 *   (Zero mean); any g_* may be NULL.  gout: upstream cotangent (device scalar) or NULL (=1). */
int64_t gpb_mll_workspace_bytes(int64_t N, int D);
int gpb_mll_forward(void* stream, int kind, int64_t N, int D, const double* X, int64_t ldx, const double* y,
                    const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                    const double* obs_stddev, const double* mean_const, double jitter, double* Sigma,
                    int64_t lds, void* ws, int64_t ws_bytes, double* value_out, double* alpha_out, int* info);
int gpb_mll_backward(void* stream, int kind, int64_t N, int D, const double* X, int64_t ldx,
                     const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                     const double* obs_stddev, double* Sigma, int64_t lds, void* ws, int64_t ws_bytes,
                     const double* alpha, const double* gout, double* g_lengthscale, double* g_variance,
                     double* g_obs_stddev, double* g_mean_const);


/* ---- collapsed_elbo (SGPR) value + analytic gradient, row-sharded ------------------------------
 * Reference: gpjax/objectives.py:321-416.  The N-sized data enter only through row-additive
 * statistics, which is what makes the path shard over GPUs:
 *   1. gpb_sgpr_stats      (per rank, local rows)  -> Paug[(M+2) x (M+2)], row stride M+2, lower
 *        triangle: sums of [A~; d^T; 1^T][A~; d^T; 1^T]^T with A~ = Lz^-1 Kzx (objectives.py:352-390)
 *   2. all-reduce(sum) of Paug over the ranks (ncclAllReduce / torch.distributed; (M+2)^2 doubles)
 *   3. gpb_sgpr_finish     (replicated)            -> ELBO (objectives.py:393-416) (+ adjoints)
 *   4. gpb_sgpr_grad_local (per rank, local rows)  -> g_Z[M,D], g_lengthscale, g_variance partials
 *   5. all-reduce(sum) of those three
 *   6. gpb_sgpr_grad_finish (replicated)           -> adds the Kzz / scalar terms, applies *gout,
 *        writes g_obs_stddev and g_mean_const.
 * need_grad of gpb_sgpr_finish / gpb_svgp_finish is a flag word: bit 0 = prepare the gradient pass; bit 1 (GPB_FINISH_DENSE_INT8) =
 * the caller vouches that Kzz is well conditioned (the precondition of gpb_sgpr_stats_raw; the Python "auto" route sets both from
 * one condition estimate, cond <= 1e3), so the dense M x M x M products of this replicated step may run as 56-bit int8
 * digit-plane products on the tcgen05 pipe (M >= 2048) instead of FP64 DMMA GEMMs.  Without the bit the arithmetic is FP64.
 * block_rows bounds the rows streamed at a time (it sizes the workspace); the local rows are cut into ceil(Nloc / block_rows)
 * blocks of equal size (multiples of 128 rows) rather than full blocks plus a ragged tail.
 * The same workspace must be passed to all calls of one evaluation.  info_out: int[2] =
 * {chol(Kzz) failure, chol(I + A A^T) failure} (0 = ok; value is NaN otherwise). */
int64_t gpb_sgpr_workspace_bytes(int64_t M, int D, int64_t block_rows);
int64_t gpb_sgpr_stats_count(int64_t M);
int gpb_sgpr_stats(void* stream, int kind, int64_t Nloc, int64_t M, int D, const double* X, int64_t ldx,
                   const double* y, const double* Z, int64_t ldz, const double* lengthscale,
                   int lengthscale_is_scalar, const double* variance, const double* obs_stddev,
                   const double* mean_const, double jitter, int64_t block_rows, void* ws, int64_t ws_bytes,
                   double* Paug);
/* Same result and contract as gpb_sgpr_stats, cheaper route: accumulates the RAW products [K_b^T|d|1]^T[K_b^T|d|1]
 * and applies Lz^-1 once to the M x M sums (forward N M^2 instead of 2 N M^2 flop).  Rounding of the raw sums is
 * amplified by cond(Kzz + jitter I) (relative error of the statistics ~ eps * sqrt(N) * cond), so use it for
 * well-conditioned Kzz only; gpb_sgpr_stats keeps the reference's whiten-first order (objectives.py:387-390). */
int gpb_sgpr_stats_raw(void* stream, int kind, int64_t Nloc, int64_t M, int D, const double* X, int64_t ldx,
                       const double* y, const double* Z, int64_t ldz, const double* lengthscale,
                       int lengthscale_is_scalar, const double* variance, const double* obs_stddev,
                       const double* mean_const, double jitter, int64_t block_rows, void* ws, int64_t ws_bytes,
                       double* Paug);
int gpb_sgpr_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz,
                    const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                    const double* obs_stddev, int64_t block_rows, void* ws, int64_t ws_bytes,
                    const double* Paug, int need_grad, double* elbo_out, int* info_out);
int gpb_sgpr_grad_local(void* stream, int kind, int64_t Nloc, int64_t M, int D, const double* X, int64_t ldx,
                        const double* y, const double* Z, int64_t ldz, const double* lengthscale,
                        int lengthscale_is_scalar, const double* variance, const double* obs_stddev,
                        const double* mean_const, int64_t block_rows, void* ws, int64_t ws_bytes, double* g_Z,
                        double* g_lengthscale, double* g_variance);
int gpb_sgpr_grad_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz,
                         const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                         const double* obs_stddev, int64_t block_rows, void* ws, int64_t ws_bytes,
                         const double* gout, double* g_Z, double* g_lengthscale, double* g_variance,
                         double* g_obs_stddev, double* g_mean_const);

/* ---- SVGP: uncollapsed minibatch ELBO (gpjax/objectives.py:241-315; VariationalGaussian.prior_kl / predict,
 * variational_families.py:169-285; KL, distributions.py:188-228; analytical Gaussian integrator,
 * integrators.py:151-158) with the analytic gradient.  The minibatch enters only through the SAME statistics as
 * the collapsed bound, so the protocol reuses the SGPR entry points:
 *   gpb_sgpr_stats (jitter = q.jitter) -> all-reduce -> gpb_svgp_finish -> gpb_sgpr_grad_local -> all-reduce
 *   -> gpb_svgp_grad_finish.
 * mu: variational_mean [M]; W: variational_root_covariance [M x M] lower triangular (strict upper ignored);
 * num_datapoints: likelihood.num_datapoints (the batch size is taken from the reduced statistics, so with G
 * ranks the effective batch is the sum of the rank batches).  g_W receives the lower triangle (upper zero). */
int gpb_svgp_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz,
                    const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                    const double* obs_stddev, const double* mean_const, const double* mu, const double* W,
                    int64_t ldw, double num_datapoints, double jitter, int64_t block_rows, void* ws,
                    int64_t ws_bytes, const double* Paug, int need_grad, double* elbo_out, int* info_out);
int gpb_svgp_grad_finish(void* stream, int kind, int64_t M, int D, const double* Z, int64_t ldz,
                         const double* lengthscale, int lengthscale_is_scalar, const double* variance,
                         const double* obs_stddev, double jitter, int64_t block_rows, void* ws, int64_t ws_bytes,
                         const double* gout, const double* W, int64_t ldw, double* g_Z, double* g_lengthscale,
                         double* g_variance, double* g_obs_stddev, double* g_mean_const, double* g_mu, double* g_W,
                         int64_t ldgw);

/* ---- exchange step of the row-sharded sparse path: all-reduce(sum) over NCCL --------------------------------------
 * Steps 2 and 5 of the protocol above.  In the reference the sum over data rows is the contraction
 * inside collapsed_elbo (gpjax/objectives.py:380-398); sharded over devices it becomes one all-reduce of Paug
 * ((M+2)^2 doubles) and one of [g_Z | g_lengthscale | g_variance].  `comm` is an ncclComm_t of the NCCL loaded in the
 * process (resolved at run time: the host framework's copy when one is mapped, else libnccl.so.2; GPB_ERR_UNSUPPORTED when
 * there is none).  gpb_allreduce_f64 is in place and only enqueues on `stream`.  The three communicator helpers let a host
 * without its own NCCL handle (the ctypes shim, a C++ trainer) build one: rank 0 calls gpb_nccl_unique_id, ships the 128
 * bytes to the other ranks by any means, and every rank calls gpb_nccl_comm_init_rank with ITS device current. */
int gpb_nccl_version(void); /* e.g. 22809; 0 = no NCCL in the process */
int gpb_nccl_unique_id(void* id_out_128_bytes);
int gpb_nccl_comm_init_rank(void** comm_out, int nranks, const void* id_128_bytes, int rank);
int gpb_nccl_comm_destroy(void* comm);
int gpb_allreduce_f64(void* comm, void* stream, double* buf, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* GPJAX_B200_H */
