// Device-primitive launcher interface.
//
// Everything float64, row-major.  All pointers are DEVICE pointers owned by the caller; every
// launcher only enqueues work on `stream` (a cudaStream_t passed as void*): no allocation, no
// synchronisation, no host reads.  The blocked algorithms in algorithms.cpp are written purely
// against this interface, so the same orchestration code can be linked against the CUDA
// implementation (primitives_cuda.cu -> the shipped libgpjax_b200.so) or against a plain C++
// host model used ONLY by the CPU test-suite (tests/hostsim/primitives_host.cpp).
#pragma once
#include <cstdint>

namespace gpb {

typedef void* stream_t;

enum KernelKind {
    KIND_RBF = 0, KIND_MATERN32 = 1, KIND_MATERN52 = 2, KIND_MATERN12 = 3,
    KIND_RATQUAD = 4,   // shape parameter: alpha
    KIND_POWEXP = 5,    // shape parameter: power
    KIND_PERIODIC = 6,  // shape parameter: period
    KIND_WHITE = 7
};
// Kinds 4..6 carry one extra scalar.  Convention on every entry point: `variance` then points to TWO
// consecutive device doubles {variance, shape} and the variance gradient output likewise to {g_variance, g_shape}.
inline bool kind_has_shape(int kind) { return kind == KIND_RATQUAD || kind == KIND_POWEXP || kind == KIND_PERIODIC; }
inline bool kind_valid(int kind) { return kind >= KIND_RBF && kind <= KIND_WHITE; }

// error codes (identical to include/gpjax_b200.h)
#ifndef GPB_OK
#define GPB_OK 0
#define GPB_ERR_INVALID (-1)      // bad argument (null pointer, negative size, unknown kind)
#define GPB_ERR_UNSUPPORTED (-2)  // e.g. D larger than the compiled maximum
#define GPB_ERR_LAUNCH (-3)       // CUDA launch error
#define GPB_ERR_WORKSPACE (-4)    // workspace too small
#endif

// Operand layouts for gemm(): element (m,k) of A / (n,k) of B lives at
//   LAYOUT_K : ptr[m*ld + k]   (K contiguous, "row-major M x K")
//   LAYOUT_MN: ptr[k*ld + m]   (M/N contiguous, "row-major K x M")
enum Layout { LAYOUT_K = 0, LAYOUT_MN = 1 };

// Output masks for gemm(): which C elements are written.  (r,c) are global coordinates
// r = mask_row0 + m, c = mask_col0 + n.
enum Mask {
    MASK_NONE = 0,
    MASK_LOWER = 1,               // r >= c
    MASK_UPPER = 2,               // r <= c
    MASK_BLOCK_STRICT_UPPER = 3,  // r / mask_nb <  c / mask_nb
    MASK_BLOCK_STRICT_LOWER = 4,  // r / mask_nb >  c / mask_nb
    // ozaki_gemm only: blocks with r / mask_nb <= c / mask_nb are live; a DIAGONAL block b is not written to C but to the dense
    // mask_nb x mask_nb matrix at C2 + b * mask_nb^2 (the exact path keeps L in the diagonal blocks of its N x N buffer and the
    // diagonal blocks of Sigma^-1 in the workspace)
    MASK_BLOCK_UPPER_DIAG_TO_C2 = 5
};

// K-range restriction exploiting triangular operands (zeros are never read).
enum KRange {
    KR_FULL = 0,
    KR_B_LOWER = 1,  // B(n,k) == 0 for k > n + kr_off   -> k <  n_tile_end + kr_off
    KR_B_UPPER = 2,  // B(n,k) == 0 for k < n + kr_off   -> k >= n_tile_begin + kr_off
    KR_A_LOWER = 3,  // A(m,k) == 0 for k > m + kr_off
    KR_A_UPPER = 4   // A(m,k) == 0 for k < m + kr_off
};

struct GemmDesc {
    int64_t M = 0, N = 0, K = 0;
    const double* A = nullptr;
    int64_t lda = 0;
    int a_layout = LAYOUT_K;
    const double* B = nullptr;
    int64_t ldb = 0;
    int b_layout = LAYOUT_K;
    double* C = nullptr;
    int64_t ldc = 0;
    double alpha = 1.0, beta = 0.0;  // C = beta*C + alpha * A * B^T   (beta == 0 never reads C)
    int mask = MASK_NONE;
    int64_t mask_row0 = 0, mask_col0 = 0, mask_nb = 1;
    int krange = KR_FULL;
    int64_t kr_off = 0;
    int batch = 1;
    int64_t strideA = 0, strideB = 0, strideC = 0;
};

// C = beta*C + alpha * A * B^T on the FP64 tensor pipe (DMMA).
int gemm(stream_t s, const GemmDesc& d);
// tuning hook (bench scripts only): 0 = 8 warps x (32x32); 8 = swizzled smem, 4 stages; anything else = default
void debug_set_gemm_variant(int v);

struct GramDesc {
    int kind = KIND_RBF;
    int64_t N = 0, M = 0;  // K is N x M
    int D = 0;
    const double* X = nullptr;  // [N, D], row stride ldx
    int64_t ldx = 0;
    const double* Z = nullptr;  // [M, D], row stride ldz
    int64_t ldz = 0;
    const double* ell = nullptr;  // device, [D] (ARD) or [1] (isotropic)
    int ell_is_scalar = 0;
    const double* variance = nullptr;  // device scalar
    double* K = nullptr;
    int64_t ldk = 0;
    int lower_only = 0;          // square case: write only tiles that touch r >= c
    double diag_add = 0.0;       // host constant added where global row == col (jitter)
    const double* diag_add_sq = nullptr;  // optional device scalar t: adds t*t on the diagonal (obs_stddev)
    int64_t row0 = 0, col0 = 0;  // global coordinates of K[0,0] (for the diagonal test)
};

// K[i,j] = variance * g(sum_d ((x_id - z_jd)/l_d)^2) (+ diagonal terms); direct-difference form.
int gram(stream_t s, const GramDesc& d);

// Cholesky of one n x n block (n <= 128), in place in the lower triangle of A (strict upper
// untouched).  Also writes inv(L) to Dinv (n x n, row stride ldd, upper part zero) and its
// transpose to DinvT.  On a non-positive / NaN pivot: sets *info = global_row0 + column + 1
// (first failure wins) and NaN-fills the block, Dinv and DinvT (JAX cholesky semantics).
// factor == 0: A already holds a lower-triangular factor; only the inverses are produced.
int potrf_leaf(stream_t s, int n, double* A, int64_t lda, double* Dinv, int64_t ldd, double* DinvT,
               int64_t lddt, int* info, int64_t global_row0, int factor);

// y = beta*y + alpha * op(A) x,  A is m x n (row stride lda); trans=0: y[m] ; trans=1: y[n].
int gemv(stream_t s, int64_t m, int64_t n, const double* A, int64_t lda, int trans, const double* x,
         double* y, double alpha, double beta);

// out[0] = sum_i log(A[i*(lda+1)])            (half log-det of L L^T)
int sum_log_diag(stream_t s, int64_t n, const double* A, int64_t lda, double* out);
// out[0] = sum_i x[i]*y[i]
int dot(stream_t s, int64_t n, const double* x, const double* y, double* out);
// out[i] = a[i] - (*c)   (c device scalar, may be null -> 0)
int sub_scalar(stream_t s, int64_t n, const double* a, const double* c, double* out);
// strided 2-D copy: dst[i*ldd + j] = src[i*lds + j]
int copy2d(stream_t s, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst,
           int64_t ldd);
// dst[j*ldd + i] = src[i*lds + j]
int transpose2d(stream_t s, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst,
                int64_t ldd);
// p[i*ld + j] = v for an rows x cols rectangle
int fill2d(stream_t s, int64_t rows, int64_t cols, double* p, int64_t ld, double v);
// zero the strict upper (uplo=2) or strict lower (uplo=1) triangle of an n x n matrix
int zero_triangle(stream_t s, int64_t n, double* A, int64_t lda, int uplo);
// mirror: A[c,r] = A[r,c] for r > c (from_lower=1) or A[r,c] = A[c,r] for r > c (from_lower=0)
int symmetrize(stream_t s, int64_t n, double* A, int64_t lda, int from_lower);
// lower triangle <- (A + A^T) / 2, strict upper untouched (jnp.linalg.cholesky symmetrises its input first)
int symmetrize_average_lower(stream_t s, int64_t n, double* A, int64_t lda);

// out[0] = -0.5*(n*log(2*pi) + 2*half_logdet[0] + quad[0]); NaN if *info != 0.
int mll_value(stream_t s, int64_t n, const double* half_logdet, const double* quad, const int* info,
              double* out);

struct MllBwdDesc {
    int kind = KIND_RBF;
    int64_t N = 0;
    int D = 0;
    int64_t nb = 256;           // block size of the Sigma^-1 storage
    const double* X = nullptr;  // [N, D]
    int64_t ldx = 0;
    const double* alpha = nullptr;  // [N]
    const double* S = nullptr;      // N x N buffer: blocks strictly above the block diagonal hold Sigma^-1
    int64_t lds = 0;
    const double* Sdiag = nullptr;  // ceil(N/nb) blocks of nb x nb (row stride nb): diagonal blocks of Sigma^-1
    const double* ell = nullptr;
    int ell_is_scalar = 0;
    const double* variance = nullptr;
    const double* obs_stddev = nullptr;
    const double* gout = nullptr;  // upstream cotangent (device scalar), may be null -> 1
    double* partials = nullptr;    // workspace, mll_bwd_partials_count(...) doubles
    double* g_ell = nullptr;       // [D] or [1]
    double* g_var = nullptr;
    double* g_obs_stddev = nullptr;
    double* g_mean = nullptr;
};
int64_t mll_bwd_partials_count(int64_t N, int D, int64_t nb);
// Streams W = 1/2 (alpha alpha^T - Sigma^-1) tile by tile against dK/dtheta recomputed from X.
int mll_bwd(stream_t s, const MllBwdDesc& d);

struct GramBwdDesc {
    int kind = KIND_RBF;
    int64_t N = 0, M = 0;
    int D = 0;
    const double* X = nullptr;
    int64_t ldx = 0;
    const double* Z = nullptr;
    int64_t ldz = 0;
    const double* ell = nullptr;
    int ell_is_scalar = 0;
    const double* variance = nullptr;
    const double* dK = nullptr;  // cotangent, N x M
    int64_t lddk = 0;
    double scale = 1.0;
    double* partials = nullptr;  // gram_bwd_partials_count(...) doubles
    // all accumulated (+=) so several blocks / the Kzz term can add into the same gradient
    double* g_ell = nullptr;  // [D] or [1]
    double* g_var = nullptr;  // [1]
    double* g_X = nullptr;    // [N, D] or null
    int64_t ldgx = 0;
    double* g_Z = nullptr;  // [M, D] or null
    int64_t ldgz = 0;
};
int64_t gram_bwd_partials_count(int64_t N, int64_t M, int D);
// g_theta += scale * <dK, dK/dtheta>, g_X / g_Z likewise; dK read once, K recomputed on the fly.
int gram_bwd(stream_t s, const GramBwdDesc& d);


// ---- small helpers used by the SGPR (collapsed_elbo) path -----------------------------------
// A = I (n x n)
int set_identity(stream_t s, int64_t n, double* A, int64_t lda);
// out[0] = sum_i x[i]
int vec_sum(stream_t s, int64_t n, const double* x, double* out);
// y[i] += alpha * x[i]
int axpy(stream_t s, int64_t n, double alpha, const double* x, double* y);
// x[i] *= (*f) (f device scalar; null -> no-op)
int scale_inplace(stream_t s, int64_t n, double* x, const double* f);
// T[r*ld + M] = y[r] - (*c) ; T[r*ld + M + 1] = 1      (c may be null)
int sgpr_aug_columns(stream_t s, int64_t rows, double* T, int64_t ld, int64_t M, const double* y,
                     const double* c);
// Unpack the (all-reduced) augmented statistics Paug [(M+2) x (M+2), lower stored]:
//   s = obs_stddev^2, Bmat = I + Phi~/s (full symmetric, ld M), psi = Paug[M,:M]/sqrt(s),
//   a1 = Paug[M+1,:M]/sqrt(s), sc = {dd, sd, n, tr(Phi~)/s, s}
int sgpr_prepare(stream_t s, int64_t M, const double* Paug, int64_t ldp, const double* obs_stddev,
                 double* Bmat, double* psi, double* a1, double* sc);
// out = 1/2 [ -n log(2 pi s) - 2 hl - (dd - wtw)/s - (n var / s - trPhi) ];  NaN if info[0] or info[1]
int sgpr_value(stream_t s, const double* sc, const double* half_logdetB, const double* wtw,
               const double* variance, const int* info, double* out);
// dPhi = 1/2 (I - Binv - v v^T / s), Phi = Bmat - I:
//   G1 = (2/s) dPhi, G2 = dPhi - Phi/2, u = v / (s sqrt s), rowsum[r] = sum_c dPhi[r,c] Phi[r,c]
int sgpr_adjoints(stream_t s, int64_t M, const double* Binv, const double* Bmat, const double* v,
                  const double* sc, double* G1, double* G2, double* u, double* rowsum);
// replicated scalar part of the gradient (dots = {psi.v, v.a1, <dPhi,Phi>}):
//   g_var += -n/(2s);  g_obs = 2 sn g_s;  g_mean = -(v.a1)/s + sd/s   with
//   g_s = -n/(2s) + (dd - psi.v)/(2 s^2) + n var/(2 s^2) - (2 <dPhi,Phi> + psi.v / s)/(2 s)
int sgpr_scalar_grads(stream_t s, const double* sc, const double* dots, const double* variance,
                      const double* obs_stddev, double* g_var, double* g_obs, double* g_mean);


// ---- SVGP (uncollapsed ELBO) finish helpers; statistics are the same Paug as SGPR -----------------------
// out[0] = sum_i log|A[i*(lda+1)]|
int sum_log_abs_diag(stream_t s, int64_t n, const double* A, int64_t lda, double* out);
// Unpack Paug WITHOUT noise scaling: Phi = A~A~^T (full symmetric, ld M), psi = A~ d, a1 = A~ 1,
//   sc = {dd, sd, B, tr(Phi), s = obs_stddev^2, coef = (num_datapoints / B) / s}
int svgp_unpack(stream_t s, int64_t M, const double* Paug, int64_t ldp, const double* obs_stddev,
                double num_datapoints, double* Phi, double* psi, double* a1, double* sc);
// dots = {u.psi, u.u, |V|_F^2, <Phi,Ttil>, half_logdet_Kzz, sum log|W_ii|}:
//   ELL = -1/2 [B log(2 pi s) + (dd - 2 u.psi + <Phi,Ttil> + B (var + jitter) - tr Phi) / s]
//   KL  = 1/2 [u.u - M - 2 sum log|W_ii| + 2 half_logdet_Kzz + |V|_F^2],  out = (N/B) ELL - KL  (NaN if info[0])
int svgp_value(stream_t s, int64_t M, const double* sc, const double* dots, const double* variance, double jitter,
               const int* info, double* out);
// G1 = coef (I - Ttil);  E = -coef sym(u psi^T) + coef/2 (PT + PT^T) - coef/2 Phi + Ttil/2 - I/2   (PT = Phi Ttil)
int svgp_adjoints(stream_t s, int64_t M, const double* Phi, const double* Ttil, const double* PT, const double* u,
                  const double* psi, const double* sc, double* G1, double* E);
// tvec = coef (psi - Phiu) - u ;  uvec = coef u
int svgp_vectors(stream_t s, int64_t M, const double* psi, const double* Phiu, const double* u, const double* sc,
                 double* tvec, double* uvec);
// H = coef * PhiV + V   (M x M, contiguous)
int svgp_h(stream_t s, int64_t M, const double* PhiV, const double* V, const double* sc, double* H);
// gW[i*ldg + i] += 1 / W[i*ldw + i]
int svgp_gw_diag(stream_t s, int64_t M, const double* W, int64_t ldw, double* gW, int64_t ldg);
// scalar gradients: g_var += -(N/B) B/(2s); g_obs = 2 sn (N/B)(-B/(2s) + Q/(2 s^2)); g_mean = coef (sd - u.a1) - sum(dF/dmu)
//   dots2 = {u.a1, sum(dF/dmu)};  Q = dd - 2 u.psi + <Phi,Ttil> + B (var + jitter) - tr Phi
int svgp_scalar_grads(stream_t s, const double* sc, const double* dots, const double* dots2, const double* variance,
                      const double* obs_stddev, double jitter, double* g_var, double* g_obs, double* g_mean);

// ---- FP64 rank-k updates on the INT8 tensor pipe (Ozaki scheme; ozaki_i8.cu) ---------------------------------------
// x_ik = scale_i * sum_p Q[i, p*k + c] * 2^(-8 (p+1)),  scale_i = 2^e_i,  -128 <= Q <= 127: `nslices` balanced radix-256 digit
// planes per entry (8 bits per plane, rounded ONCE to the last plane), plane p stored at columns [p*k, (p+1)*k) of the int8
// matrix Q (row stride ldq >= nslices*k bytes, multiple of 16; Q 16-byte aligned).  |x_ik| <= 0.494 scale_i.  Exact (no rounding)
// for entries whose last mantissa bit lies above 2^(-8 nslices) scale_i; a NaN/Inf row gets scale = NaN.
// Plane stride `kplane` >= k (columns [k, kplane) of every plane are written as zero digits, so a ragged K can be padded
// to the multiple of 128 the product kernel needs).
// nslices_dev (optional DEVICE word, the guard of ozaki_choose_planes): when given, min(nslices, *nslices_dev) planes are extracted
// and the ONE rounding happens at the last plane the product will use -- a prefix of a longer digit string would be a truncation,
// whose error has a non-zero mean (balanced digits in [-128, 127] average -1/2) that adds up coherently over K and over the
// block steps of a factorisation.
int ozaki_slice(stream_t s, int64_t rows, int64_t k, int64_t kplane, const double* X, int64_t ldx, int nslices, int8_t* Q,
                int64_t ldq, double* scale, const int* nslices_dev = nullptr);
// Same digits for the COLUMNS of a rows x cols block (contraction over the rows): Qt[c, p*kplane + r], scale[c] from the column
// maxima; rows [rows, kplane) are zero digits (kplane multiple of 128); colmax_scratch: cols doubles.
int ozaki_slice_t(stream_t s, int64_t rows, int64_t cols, int64_t kplane, const double* X, int64_t ldx, int nslices, int8_t* Qt,
                  int64_t ldq, double* scale, double* colmax_scratch);
// out_w[c] += sum_r w[r] X[r,c], out_1[c] += sum_r X[r,c] for a rows x cols block (deterministic two-stage reduction);
// scratch: 2 * ceil(rows / 1024) * cols doubles
int col_weighted_sums(stream_t s, int64_t rows, int64_t cols, const double* X, int64_t ldx, const double* w, double* scratch,
                      double* out_w, double* out_1);
constexpr int OZ_DIGIT_BITS = 8;             // radix 256
constexpr int OZ_DIGIT_SQ_MAX = 128 * 128;   // largest digit-pair product: int32 headroom is nslices * K * 2^14 < 2^31
constexpr int OZ_PLANES_MAX = 7;             // 56 bits >= the fp64 mantissa (and 8 * 7 + 7 < 63: the fixed-point form fits int64)
// Fused Gram tile -> digit planes (SGPR / SVGP streamed passes): the K_b block is never written as fp64.  Scales are FIXED from
// the bound |k(x, z)| <= variance instead of measured maxima (no pass over the block is needed for them):
//   cols_mode 0 (pass 2): Q[r, p*kplane + c] = digit p of the row [K_b | d | 1 | 0..] (c < kplane), d_r = y_r - mean;
//                         scale[r] = 2^e, e from max(|variance|, 1, |d_r|)
//   cols_mode 1 (pass 1): Qt[c, p*kplane + r] = digit p of K_b[r, c] (c < M; rows >= N up to kplane are zero digits),
//                         scale[c] = 2^e(|variance|); part[(t*2 + 0)*(M+2) + c] = sum over the 64 rows of tile row t of
//                         d_r * v_rc, part[(t*2 + 1)*(M+2) + c] = sum v_rc, for the M + 2 columns v = [K_b | d | 1]
//                         (reduced in tile-row order by col_partials_reduce: deterministic)
struct GramDigitsDesc {
    GramDesc g;                 // kind, N (rows of the block), M, D, X, Z, ell, variance; K / ldk unused
    int cols_mode = 0;
    int nslices = OZ_PLANES_MAX;
    int8_t* Q = nullptr;
    int64_t ldq = 0, kplane = 0;   // kplane: multiple of 128, >= M + 2 (mode 0) / >= N (mode 1)
    double* scale = nullptr;
    const double* y = nullptr;           // [N]
    const double* mean_const = nullptr;  // device scalar or null (zero mean)
    double* part = nullptr;              // mode 1: 2 * gram_digits_tile_rows(kplane) * (M + 2) doubles
};
int gram_digits(stream_t s, const GramDigitsDesc& d);
inline int64_t gram_digits_tile_rows(int64_t kplane) { return (kplane + 63) / 64; }
// out_w[c] += sum_t part[(t*2+0)*cols + c], out_1[c] += sum_t part[(t*2+1)*cols + c]   (t ascending)
int col_partials_reduce(stream_t s, int64_t chunks, int64_t cols, const double* part, double* out_w, double* out_1);
struct OzakiGemmDesc {
    int64_t M = 0, N = 0, K = 0;  // K = digits per plane (multiple of 128)
    int nslices = 6;              // digit planes present in Qa / Qb (and used, unless nslices_dev overrides it)
    const int* nslices_dev = nullptr;  // optional DEVICE word: planes to use, 1..nslices (the guard of ozaki_choose_planes);
                                       // read by the kernel, so the host never synchronises on the decision
    const int8_t* Qa = nullptr;   // [M, nslices*K]
    int64_t ldqa = 0;
    const double* sa = nullptr;   // [M] row scales
    const int8_t* Qb = nullptr;   // [N, nslices*K]
    int64_t ldqb = 0;
    const double* sb = nullptr;   // [N]
    double* C = nullptr;          // C += alpha * A B^T (all digit pairs of order p+q < nslices); beta0: C = alpha * A B^T
    int beta0 = 0;
    int64_t plane_stride = 0;     // digits between consecutive planes of one row (0 -> K): lets one launch cover a K sub-range
    int64_t ldc = 0;
    double alpha = 1.0;
    int mask = MASK_NONE;         // MASK_NONE, MASK_LOWER, MASK_BLOCK_STRICT_UPPER (as GemmDesc::mask) or MASK_BLOCK_UPPER_DIAG_TO_C2
    int64_t mask_row0 = 0, mask_col0 = 0, mask_nb = 1;
    double* C2 = nullptr;         // MASK_BLOCK_UPPER_DIAG_TO_C2: the diagonal blocks (mask_nb multiple of 128)
    int krange = KR_FULL;         // triangular operand: its structural zeros are never multiplied (K-blocks of 128 digits skipped)
    int64_t kr_off = 0;
    int max_ctas = 0;             // > 0: leave SMs free for a concurrent latency-bound chain on another stream (look-ahead)
};
int ozaki_gemm(stream_t s, const OzakiGemmDesc& d);
// SMs of the current device (148 on B200), for callers that size OzakiGemmDesc::max_ctas
int device_sm_count();
// krange / MASK_BLOCK_UPPER_DIAG_TO_C2 / max_ctas need the CTA-pair kernel (the default; GPB_OZ_KERNEL=1|2 select older variants)
bool ozaki_supports_extensions();
// Digit planes of the int8 trailing updates of an N x N covariance factorisation, decided ON THE DEVICE (no host read):
//   requested in 1..7        -> planes_out[0] = requested;
//   requested == OZ_AUTO (-1)-> 6 if the hyper-parameters PROVE cond(Sigma) <= OZ_AUTO_COND_LIMIT, else 7, with the bound
//                               cond(K + s I) <= (N * variance + s) / s,  s = obs_stddev^2 + jitter   (|k(x,y)| <= variance =>
//                               lambda_max(K) <= N variance by Gershgorin; lambda_min(Sigma) >= s);
//                               variance == nullptr (a bare matrix, nothing known about it) -> 7.
// Calibration (profiles/r02_cond_sweep_radix256.jsonl, N = 8192, 24 cells, cond 1e3 .. 3e8): 7 planes (56 bits) sit at the FP64 DMMA
// path's own error level at every cond; 6 planes (48 bits + the equal-plane term) carry at most 1.4e-15 * BOUND relative error in the
// most sensitive output (the mean-constant gradient 1^T Sigma^-1 (y - m) of a very smooth kernel), i.e. <= 2.8e-9 under the guard
// (contract 1e-8); BOUND over-estimates cond(Sigma) by 3x in that cell and by 10-300x in the others.
constexpr int OZ_AUTO = -1;
constexpr int OZ_AUTO_PLANES_LO = 6, OZ_AUTO_PLANES_HI = 7;
constexpr double OZ_AUTO_COND_LIMIT = 2e6;
int ozaki_choose_planes(stream_t s, int requested, int64_t N, const double* variance, const double* obs_stddev, double jitter,
                        int* planes_out);
// the same rule on the host, for reporting (bench.py) and tests; never used to steer a launch
int ozaki_auto_planes_host(int64_t N, double variance, double obs_stddev, double jitter);
// C[m,n] (int32) = A[m,k] B[n,k]^T for int8 operands (k multiple of 128, lda/ldb multiples of 16): the raw tcgen05 product
int igemm_i8(stream_t s, int64_t m, int64_t n, int64_t k, const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb,
             int32_t* C, int64_t ldc);
// false when the driver cannot encode TMA tensor maps (then the blocked algorithms stay on the DMMA path)
bool ozaki_available();

// ---- intra-call concurrency (lookahead) ---------------------------------------------------------------
// side_stream: a lazily created, higher-priority helper stream of the current device (index 0..3).
// stream_fork(from, to): everything enqueued on `to` afterwards waits for what is on `from` now
// (event record + wait; no host synchronisation).  Work put on a side stream must be joined back
// into the caller's stream (stream_fork(side, main)) before the entry point returns.
stream_t side_stream(stream_t main, int idx);
int stream_fork(stream_t from, stream_t to);

// ---- optional measurement hooks (bench.py roofline): off by default, zero cost when off ----------
// enable != 0: every gemm() launch is bracketed by CUDA events on its own stream and every kernel
// launch of the library is counted.  profile_read synchronises the recorded events.
void profile_reset(int enable);
int profile_read(double* gemm_ms, int64_t* gemm_launches, int64_t* all_launches);
void profile_count_launch();
bool profile_enabled();
void profile_gemm_begin(stream_t s);
void profile_gemm_end(stream_t s);
// same bracket for the Ozaki int8 kernel (profile_gemm_end closes it); int8_ops = algorithmic int8 operations of the launch
void profile_ozaki_begin(stream_t s, double int8_ops);
int profile_read_ozaki(double* ms, int64_t* launches, double* int8_ops);

// maximum input dimension D the compiled kernels support
int max_input_dim();

}  // namespace gpb
