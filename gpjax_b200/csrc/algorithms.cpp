// Blocked right-looking algorithms (host orchestration only -- see algorithms.h).
//
// Storage convention for the exact-GP path (one N x N row-major buffer A, block size NB = ws.nb: 1024 below 12,288 rows, 2048
// from there on -- algorithms.h):
//   * lower triangle incl. the diagonal blocks : Sigma, then its Cholesky factor L (in place)
//   * blocks strictly above the block diagonal : W = L^-T (trtri), then Sigma^-1 = W W^T (lauum)
//   * ws.Dinv / ws.DinvT                       : inverses of the diagonal blocks of L (and transposes)
//   * ws.Sdiag                                 : diagonal blocks of Sigma^-1
// so a single N x N buffer carries forward AND backward (N = 100k -> 80 GB of the 180 GB HBM) and
// L survives the backward pass.  Every O(N^3) step is a rank-NB update C += A B^T (int8 digit planes on tcgen05 for updates
// with >= 2048 rows -- csrc/ozaki_i8.cu -- FP64 DMMA GEMM otherwise); the panel x inverse-block products next to them take the
// same pipe (oz_tri_product):
//   potrf : panel  X = P * inv(L_kk)^T,  trailing  A22 -= X X^T          (N^3/3)
//   trtri : W[:,k] = -Acc * inv(L_kk)^T, Acc[:, k+1:] += W[:,k] L[k+1:,k]^T   (N^3/3)
//   lauum : S[:k,:k] += W[:,k] W[:,k]^T, S[:,k] = W[:,k] inv(L_kk)       (N^3/3)
// All operands are addressed "K-contiguous" (C = A * B^T with row-major A and B), which is what the
// row-major lower/upper split gives for free.
#include "algorithms.h"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace gpb {

#define GPB_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != GPB_OK) return rc__; \
    } while (0)

// -------------------------------------------------------------------------------------------
// block size rule (algorithms.h)
// -------------------------------------------------------------------------------------------
int64_t block_size_for(int64_t ws_n) {
#ifdef GPB_NB
    (void)ws_n;
    return GPB_NB;
#else
    static const int64_t large = [] {
        const char* e = std::getenv("GPB_NB_LARGE");
        const long long v = e ? std::atoll(e) : NB_LARGE;
        return (v >= 128 && v <= 8192 && v % 128 == 0) ? (int64_t)v : NB_LARGE;
    }();
    static const int64_t min_rows = [] {
        const char* e = std::getenv("GPB_NB_LARGE_MIN_ROWS");
        const long long v = e ? std::atoll(e) : NB_LARGE_MIN_ROWS;
        return v > 0 ? (int64_t)v : NB_LARGE_MIN_ROWS;
    }();
    return ws_n >= min_rows ? large : NB_SMALL;
#endif
}

// -------------------------------------------------------------------------------------------
// Ozaki switch
// -------------------------------------------------------------------------------------------
namespace {
constexpr int OZ_UNSET = -100;
std::atomic<int> g_oz_slices{OZ_UNSET};
int clamp_slices(int v) { return (v == OZ_AUTO || (v >= OZ_MIN_SLICES && v <= OZ_MAX_SLICES)) ? v : 0; }
}  // namespace
void set_ozaki_slices(int nslices) { g_oz_slices.store(clamp_slices(nslices), std::memory_order_relaxed); }
int get_ozaki_slices() {
    int v = g_oz_slices.load(std::memory_order_relaxed);
    if (v == OZ_UNSET) {
        const char* e = std::getenv("GPB_OZAKI");
        v = (!e || !std::strcmp(e, "auto")) ? clamp_slices(GPB_OZ_DEFAULT) : clamp_slices(std::atoi(e));
        g_oz_slices.store(v, std::memory_order_relaxed);
    }
    return v;
}
// does a rank-NB update with `rows` output rows run on the int8 pipe?  (plane count: ws.oz_planes, decided on the device)
static bool oz_on(const FactorWs& ws, int64_t rows) {
    return get_ozaki_slices() != 0 && ws.oz_q && ws.oz_planes && rows >= OZ_MIN_ROWS && ws.nb % 128 == 0 && ozaki_available();
}
// Panel x inverse-block products (X = P inv(L_kk)^T and friends: M x NB x NB with a triangular NB x NB operand) on the int8 pipe as
// well: GPB_OZ_PANELS=0 keeps them on the DMMA pipe.  Same predicate as the rank-NB update next to each of them (oz_on), full
// blocks only; the inverse block is sliced like any other operand (power-of-two row scales), its structural zeros are skipped by
// the kernel's K-range.
static bool oz_panels_on(const FactorWs& ws, int64_t rows, int64_t nbk) {
    static const bool on = [] {
        const char* e = std::getenv("GPB_OZ_PANELS");
        return !e || std::atoi(e) != 0;
    }();
    return on && nbk == ws.nb && oz_on(ws, rows) && ozaki_supports_extensions();
}
// C[0:M, 0:NB] = alpha * A T^T with A's digit planes at (qa, sa) and the triangular block T (NB x NB, row stride NB) sliced into
// (qt, st): NB rows of scratch planes.  krange says where T's zeros are (KR_B_*), or, with swap, T is the LEFT operand:
// C[0:NB, 0:M] = alpha * T A^T (KR_A_*).
static int oz_tri_product(stream_t s, const FactorWs& ws, int64_t M, const int8_t* qa, const double* sa, const double* T, int8_t* qt,
                          double* st, double* C, int64_t ldc, double alpha, int krange, bool swap) {
    const int64_t NB = ws.nb;
    const int64_t ldq = OZ_MAX_SLICES * NB;
    GPB_TRY(ozaki_slice(s, NB, NB, NB, T, NB, OZ_MAX_SLICES, qt, ldq, st, ws.oz_planes));
    OzakiGemmDesc g;
    g.K = NB; g.nslices = OZ_MAX_SLICES; g.nslices_dev = ws.oz_planes;
    if (!swap) {
        g.M = M; g.N = NB; g.Qa = qa; g.sa = sa; g.Qb = qt; g.sb = st;
    } else {
        g.M = NB; g.N = M; g.Qa = qt; g.sa = st; g.Qb = qa; g.sb = sa;
    }
    g.ldqa = ldq; g.ldqb = ldq;
    g.C = C; g.ldc = ldc; g.alpha = alpha; g.beta0 = 1; g.krange = krange;
    return ozaki_gemm(s, g);
}

int factor_set_planes(stream_t s, const FactorWs& ws, int64_t N, const double* variance, const double* obs_stddev, double jitter) {
    const int req = get_ozaki_slices();
    if (req == 0 || !ws.oz_planes || !ozaki_available()) return GPB_OK;
    return ozaki_choose_planes(s, req, N, variance, obs_stddev, jitter, ws.oz_planes);
}

// -------------------------------------------------------------------------------------------
// workspace
// -------------------------------------------------------------------------------------------
namespace {
struct WsLayout {
    int64_t off_dinv, off_dinvt, off_sdiag, off_panel, off_panel2, off_small, off_vec, off_scal, off_part, total;
    int64_t off_ozq, off_ozq2, off_ozs, off_ozs2, off_ozp;
    bool with_oz;
    int64_t partials_count;
};
WsLayout ws_layout(int64_t N, int D, int with_potri) {
    WsLayout L;
    const int64_t NB = block_size_for(N);
    const int64_t blk = nblocks(N, NB) * NB * NB;
    int64_t o = 0;
    auto take = [&](int64_t n) {
        int64_t r = o;
        o += align_up(n, 32);  // 256-byte granularity
        return r;
    };
    L.off_dinv = take(blk);
    L.off_dinvt = take(blk);
    L.off_sdiag = take(with_potri ? blk : 0);
    L.off_panel = take(align_up(N, NB) * NB);
    L.off_panel2 = take(align_up(N, NB) * NB);
    L.off_small = take(4 * NB * NB);
    L.off_vec = take(4 * align_up(N, NB));
    L.off_scal = take(16);
    L.partials_count = with_potri ? mll_bwd_partials_count(N, D > 0 ? D : 1, NB) : 0;
    L.off_part = take(L.partials_count);
    // Ozaki digit buffers: OZ_MAX_SLICES * NB bytes per row (NB is a multiple of 8)
    L.with_oz = N >= OZ_MIN_ROWS + NB;
    const int64_t qd = L.with_oz ? align_up(N, NB) * (OZ_MAX_SLICES * NB / 8) : 0;
    L.off_ozq = take(qd);
    L.off_ozq2 = take(qd);
    L.off_ozs = take(L.with_oz ? align_up(N, NB) : 0);
    L.off_ozs2 = take(L.with_oz ? align_up(N, NB) : 0);
    L.off_ozp = take(L.with_oz ? 32 : 0);
    L.total = o;
    return L;
}
}  // namespace

int64_t factor_ws_bytes(int64_t N, int D, int with_potri) {
    if (N < 0) return 0;
    return ws_layout(N, D, with_potri).total * (int64_t)sizeof(double);
}

int factor_ws_carve(void* buf, int64_t bytes, int64_t N, int D, int with_potri, FactorWs* ws) {
    if (!buf || !ws) return GPB_ERR_INVALID;
    WsLayout L = ws_layout(N, D, with_potri);
    if (bytes < L.total * (int64_t)sizeof(double)) return GPB_ERR_WORKSPACE;
    double* b = static_cast<double*>(buf);
    ws->nb = block_size_for(N);
    ws->Dinv = b + L.off_dinv;
    ws->DinvT = b + L.off_dinvt;
    ws->Sdiag = with_potri ? b + L.off_sdiag : nullptr;
    ws->panel = b + L.off_panel;
    ws->panel2 = b + L.off_panel2;
    ws->small = b + L.off_small;
    ws->vec = b + L.off_vec;
    ws->scal = b + L.off_scal;
    ws->partials = with_potri ? b + L.off_part : nullptr;
    ws->partials_count = L.partials_count;
    ws->oz_q = L.with_oz ? reinterpret_cast<int8_t*>(b + L.off_ozq) : nullptr;
    ws->oz_q2 = L.with_oz ? reinterpret_cast<int8_t*>(b + L.off_ozq2) : nullptr;
    ws->oz_scale = L.with_oz ? b + L.off_ozs : nullptr;
    ws->oz_scale2 = L.with_oz ? b + L.off_ozs2 : nullptr;
    ws->oz_planes = L.with_oz ? reinterpret_cast<int*>(b + L.off_ozp) : nullptr;
    return GPB_OK;
}

// -------------------------------------------------------------------------------------------
// diagonal block: factor (optional) + explicit inverse, n <= NB, by recursive halving down to the
// single-CTA leaves (n <= LEAFN); the off-diagonal parts are small DMMA GEMMs:
//   L21 = A21 inv(L11)^T,  A22 -= L21 L21^T,  inv(L)21 = -inv(L22) L21 inv(L11)
// D / DT are the inverse and its transpose, row stride NB.  `small` provides 2*h*h doubles of
// scratch per recursion level (h = 256, 128), laid out back to back.
// -------------------------------------------------------------------------------------------
static int diag_block(stream_t s, int n, double* Ablk, int64_t lda, double* D, double* DT, double* small,
                      int* info, int64_t row0, int factor, int64_t NB) {
    if (n <= LEAFN) return potrf_leaf(s, n, Ablk, lda, D, NB, DT, NB, info, row0, factor);
    int h = (int)LEAFN;
    while (2 * h < n) h *= 2;  // largest power-of-two multiple of LEAFN strictly below n
    const int n1 = h, n2 = n - h;
    double* t21 = small;                     // [n2 x n1], ld h : L21
    double* tt = small + (int64_t)h * h;     // [n2 x n1], ld h : L21 * inv(L11)
    double* next_small = small + 2 * (int64_t)h * h;
    double* A21 = Ablk + (int64_t)n1 * lda;
    double* A22 = A21 + n1;
    double* D22 = D + (int64_t)n1 * NB + n1;
    double* DT22 = DT + (int64_t)n1 * NB + n1;
    GPB_TRY(diag_block(s, n1, Ablk, lda, D, DT, next_small, info, row0, factor, NB));
    GemmDesc g;
    if (factor) {
        // L21 = A21 * inv(L11)^T
        g = GemmDesc();
        g.M = n2; g.N = n1; g.K = n1;
        g.A = A21; g.lda = lda; g.B = D; g.ldb = NB; g.C = t21; g.ldc = h;
        g.krange = KR_B_LOWER;
        GPB_TRY(gemm(s, g));
        GPB_TRY(copy2d(s, n2, n1, t21, h, A21, lda));
        // A22 -= L21 L21^T (lower)
        g = GemmDesc();
        g.M = n2; g.N = n2; g.K = n1;
        g.A = t21; g.lda = h; g.B = t21; g.ldb = h; g.C = A22; g.ldc = lda;
        g.alpha = -1.0; g.beta = 1.0; g.mask = MASK_LOWER;
        GPB_TRY(gemm(s, g));
    } else {
        GPB_TRY(copy2d(s, n2, n1, A21, lda, t21, h));
    }
    GPB_TRY(diag_block(s, n2, A22, lda, D22, DT22, next_small, info, row0 + n1, factor, NB));
    // tt = L21 * inv(L11)          (B operand = inv(L11)^T, upper triangular)
    g = GemmDesc();
    g.M = n2; g.N = n1; g.K = n1;
    g.A = t21; g.lda = h; g.B = DT; g.ldb = NB; g.C = tt; g.ldc = h;
    g.krange = KR_B_UPPER;
    GPB_TRY(gemm(s, g));
    // inv(L)21 = -inv(L22) * tt    (B operand tt is N-contiguous)
    g = GemmDesc();
    g.M = n2; g.N = n1; g.K = n2;
    g.A = D22; g.lda = NB; g.B = tt; g.ldb = h; g.b_layout = LAYOUT_MN;
    g.C = D + (int64_t)n1 * NB; g.ldc = NB; g.alpha = -1.0;
    g.krange = KR_A_LOWER;
    GPB_TRY(gemm(s, g));
    GPB_TRY(fill2d(s, n1, n2, D + n1, NB, 0.0));
    GPB_TRY(transpose2d(s, n2, n1, D + (int64_t)n1 * NB, NB, DT + n1, NB));
    GPB_TRY(fill2d(s, n2, n1, DT + (int64_t)n1 * NB, NB, 0.0));
    return GPB_OK;
}

// SMs left to the look-ahead chain during the big int8 update of a potrf step (GPB_LOOKAHEAD_FREE_SMS, default 4; 0 = none)
constexpr int64_t LOOKAHEAD_MIN_ROWS = 12288;
static int lookahead_ctas() {
    static const int free_sms = [] {
        const char* e = std::getenv("GPB_LOOKAHEAD_FREE_SMS");
        const int v = e ? std::atoi(e) : 4;
        return v < 0 ? 0 : (v > 64 ? 64 : v);
    }();
    if (free_sms == 0) return 0;
    return device_sm_count() - free_sms;
}

// Right-looking blocked Cholesky with one-step LOOKAHEAD: the trailing update of step k is split into
// the next block column (U1) and the rest (U2); as soon as U1 is done the latency-bound chain of step
// k+1 (diagonal-block factorisation + inverse, panel X = P inv(L)^T) runs on a high-priority side stream
// while U2 -- >95 % of the step's flops -- keeps the SMs busy on the caller's stream.
static int potrf_panel(stream_t s, int64_t N, double* A, int64_t lda, const FactorWs& ws, int* info, int64_t k,
                       double* panel, int8_t* oz_q, double* oz_scale) {
    const int64_t NB = ws.nb;
    const int64_t j0 = k * NB;
    const int nbk = (int)((N - j0) < NB ? (N - j0) : NB);
    double* Dk = ws.Dinv + k * NB * NB;
    GPB_TRY(diag_block(s, nbk, A + j0 * lda + j0, lda, Dk, ws.DinvT + k * NB * NB, ws.small, info, j0, 1, NB));
    const int64_t rows = N - j0 - nbk;
    if (rows <= 0) return GPB_OK;
    double* P = A + (j0 + nbk) * lda + j0;
    if (oz_panels_on(ws, rows, nbk)) {  // X = P * inv(L_kk)^T on the int8 pipe: planes of P in rows [0, rows) of this step's digit
        const int64_t ldq = OZ_MAX_SLICES * NB;  // buffer (X's planes replace them below), planes of inv(L_kk) in the NB rows after
        GPB_TRY(ozaki_slice(s, rows, NB, NB, P, lda, OZ_MAX_SLICES, oz_q, ldq, oz_scale, ws.oz_planes));
        GPB_TRY(oz_tri_product(s, ws, rows, oz_q, oz_scale, Dk, oz_q + rows * ldq, oz_scale + rows, panel, NB, 1.0, KR_B_LOWER, false));
    } else {
        GemmDesc g;  // X = P * inv(L_kk)^T -> contiguous copy, then back in place
        g.M = rows; g.N = nbk; g.K = nbk;
        g.A = P; g.lda = lda; g.B = Dk; g.ldb = NB; g.C = panel; g.ldc = NB;
        g.krange = KR_B_LOWER;
        GPB_TRY(gemm(s, g));
    }
    GPB_TRY(copy2d(s, rows, nbk, panel, NB, P, lda));
    // Ozaki path: digit planes of the panel, extracted on the stream that produced it (the side stream under lookahead)
    if (nbk == NB && oz_on(ws, rows))  // ws.oz_planes[0] planes are extracted (rounded at the last one) and used by the product
        GPB_TRY(ozaki_slice(s, rows, NB, NB, panel, NB, OZ_MAX_SLICES, oz_q, OZ_MAX_SLICES * NB, oz_scale, ws.oz_planes));
    return GPB_OK;
}

int potrf_lower(stream_t s, int64_t N, double* A, int64_t lda, const FactorWs& ws, int* info) {
    if (N < 0 || (N > 0 && !A)) return GPB_ERR_INVALID;
    const int64_t NB = ws.nb;
    const int64_t nblk = nblocks(N, NB);
    if (nblk == 0) return GPB_OK;
    stream_t side = side_stream(s, 0);
    double* pcur = ws.panel;
    double* pnext = ws.panel2;
    int8_t* qcur = ws.oz_q;
    int8_t* qnext = ws.oz_q2;
    double* scur = ws.oz_scale;
    double* snext = ws.oz_scale2;
    const int64_t ldq = OZ_MAX_SLICES * NB;
    GPB_TRY(potrf_panel(s, N, A, lda, ws, info, 0, pcur, qcur, scur));
    for (int64_t k = 0; k + 1 < nblk; ++k) {
        const int64_t j1 = (k + 1) * NB;        // first row/column of the trailing matrix
        const int64_t rows = N - j1;             // pcur holds X_k: rows x NB
        const int64_t nb1 = rows < NB ? rows : NB;
        const bool planes = oz_on(ws, rows);  // same predicate potrf_panel used when it sliced X_k
        // U1: next block column, A[j1:, j1:j1+nb1] -= X X[0:nb1]^T (lower part)
        if (planes) {
            OzakiGemmDesc u;
            u.M = rows; u.N = nb1; u.K = NB; u.nslices = OZ_MAX_SLICES; u.nslices_dev = ws.oz_planes;
            u.Qa = qcur; u.ldqa = ldq; u.sa = scur; u.Qb = qcur; u.ldqb = ldq; u.sb = scur;
            u.C = A + j1 * lda + j1; u.ldc = lda; u.alpha = -1.0; u.mask = MASK_LOWER;
            GPB_TRY(ozaki_gemm(s, u));
        } else {
            GemmDesc u;
            u.M = rows; u.N = nb1; u.K = NB;
            u.A = pcur; u.lda = NB; u.B = pcur; u.ldb = NB;
            u.C = A + j1 * lda + j1; u.ldc = lda;
            u.alpha = -1.0; u.beta = 1.0; u.mask = MASK_LOWER;
            GPB_TRY(gemm(s, u));
        }
        const int64_t rest = rows - nb1;
        if (rest > 0) {
            GPB_TRY(stream_fork(s, side));
            GPB_TRY(potrf_panel(side, N, A, lda, ws, info, k + 1, pnext, qnext, snext));
            // U2: A[j1+nb1:, j1+nb1:] -= X[nb1:] X[nb1:]^T (lower part)
            if (planes) {
                OzakiGemmDesc v;
                v.M = rest; v.N = rest; v.K = NB; v.nslices = OZ_MAX_SLICES; v.nslices_dev = ws.oz_planes;
                v.Qa = qcur + nb1 * ldq; v.ldqa = ldq; v.sa = scur + nb1;
                v.Qb = v.Qa; v.ldqb = ldq; v.sb = v.sa;
                v.C = A + (j1 + nb1) * lda + (j1 + nb1); v.ldc = lda; v.alpha = -1.0; v.mask = MASK_LOWER;
                // A persistent launch on every SM would leave nothing for the look-ahead chain on the side stream (its kernels
                // cannot co-reside with a 198 KB / 64 K-register CTA), i.e. no overlap at all: keep a few SMs free while the update
                // is long enough to cover the chain (diagonal-block recursion ~2 ms).
                if (rest >= LOOKAHEAD_MIN_ROWS && ozaki_supports_extensions()) v.max_ctas = lookahead_ctas();
                GPB_TRY(ozaki_gemm(s, v));
            } else {
                GemmDesc v;
                v.M = rest; v.N = rest; v.K = NB;
                v.A = pcur + nb1 * NB; v.lda = NB; v.B = pcur + nb1 * NB; v.ldb = NB;
                v.C = A + (j1 + nb1) * lda + (j1 + nb1); v.ldc = lda;
                v.alpha = -1.0; v.beta = 1.0; v.mask = MASK_LOWER;
                GPB_TRY(gemm(s, v));
            }
            GPB_TRY(stream_fork(side, s));
        } else {
            GPB_TRY(potrf_panel(s, N, A, lda, ws, info, k + 1, pnext, qnext, snext));
        }
        double* t = pcur; pcur = pnext; pnext = t;
        int8_t* tq = qcur; qcur = qnext; qnext = tq;
        double* ts = scur; scur = snext; snext = ts;
    }
    return GPB_OK;
}

int diag_inverses(stream_t s, int64_t N, const double* L, int64_t lda, const FactorWs& ws) {
    const int64_t NB = ws.nb;
    const int64_t nblk = nblocks(N, NB);
    for (int64_t k = 0; k < nblk; ++k) {
        const int64_t j0 = k * NB;
        const int nbk = (int)((N - j0) < NB ? (N - j0) : NB);
        GPB_TRY(diag_block(s, nbk, const_cast<double*>(L) + j0 * lda + j0, lda, ws.Dinv + k * NB * NB,
                           ws.DinvT + k * NB * NB, ws.small, nullptr, j0, 0, NB));
    }
    return GPB_OK;
}

int trsv_lower(stream_t s, int64_t N, const double* L, int64_t lda, const FactorWs& ws, double* x, int trans) {
    const int64_t NB = ws.nb;
    const int64_t nblk = nblocks(N, NB);
    double* t = ws.vec + 2 * align_up(N, NB);  // [NB] scratch
    if (!trans) {
        for (int64_t k = 0; k < nblk; ++k) {
            const int64_t j0 = k * NB;
            const int64_t nbk = (N - j0) < NB ? (N - j0) : NB;
            GPB_TRY(gemv(s, nbk, nbk, ws.Dinv + k * NB * NB, NB, 0, x + j0, t, 1.0, 0.0));
            const int64_t rows = N - j0 - nbk;
            if (rows > 0) GPB_TRY(gemv(s, rows, nbk, L + (j0 + nbk) * lda + j0, lda, 0, t, x + j0 + nbk, -1.0, 1.0));
            GPB_TRY(copy2d(s, 1, nbk, t, NB, x + j0, NB));
        }
    } else {
        for (int64_t k = nblk - 1; k >= 0; --k) {
            const int64_t j0 = k * NB;
            const int64_t nbk = (N - j0) < NB ? (N - j0) : NB;
            GPB_TRY(gemv(s, nbk, nbk, ws.DinvT + k * NB * NB, NB, 0, x + j0, t, 1.0, 0.0));
            if (j0 > 0) GPB_TRY(gemv(s, nbk, j0, L + j0 * lda, lda, 1, t, x, -1.0, 1.0));
            GPB_TRY(copy2d(s, 1, nbk, t, NB, x + j0, NB));
        }
    }
    return GPB_OK;
}

int trsm_lower_left(stream_t s, int64_t N, int64_t T, const double* L, int64_t lda, const FactorWs& ws, double* B,
                    int64_t ldb, int trans) {
    if (N <= 0 || T <= 0) return GPB_OK;
    const int64_t NB = ws.nb;
    const int64_t nblk = nblocks(N, NB);
    double* Xk = ws.panel;  // [NB x T], row stride T
    if (!trans) {
        for (int64_t k = 0; k < nblk; ++k) {
            const int64_t j0 = k * NB;
            const int64_t nbk = (N - j0) < NB ? (N - j0) : NB;
            GemmDesc g;  // X_k = Dinv_k * B_k
            g.M = nbk; g.N = T; g.K = nbk;
            g.A = ws.Dinv + k * NB * NB; g.lda = NB;
            g.B = B + j0 * ldb; g.ldb = ldb; g.b_layout = LAYOUT_MN;
            g.C = Xk; g.ldc = T;
            g.krange = KR_A_LOWER;
            GPB_TRY(gemm(s, g));
            GPB_TRY(copy2d(s, nbk, T, Xk, T, B + j0 * ldb, ldb));
            const int64_t rows = N - j0 - nbk;
            if (rows > 0) {
                GemmDesc u;  // B[j0+nbk:, :] -= L[j0+nbk:, j0:j0+nbk] * X_k
                u.M = rows; u.N = T; u.K = nbk;
                u.A = L + (j0 + nbk) * lda + j0; u.lda = lda;
                u.B = Xk; u.ldb = T; u.b_layout = LAYOUT_MN;
                u.C = B + (j0 + nbk) * ldb; u.ldc = ldb; u.alpha = -1.0; u.beta = 1.0;
                GPB_TRY(gemm(s, u));
            }
        }
    } else {
        for (int64_t k = nblk - 1; k >= 0; --k) {
            const int64_t j0 = k * NB;
            const int64_t nbk = (N - j0) < NB ? (N - j0) : NB;
            GemmDesc g;  // X_k = Dinv_k^T * B_k
            g.M = nbk; g.N = T; g.K = nbk;
            g.A = ws.DinvT + k * NB * NB; g.lda = NB;
            g.B = B + j0 * ldb; g.ldb = ldb; g.b_layout = LAYOUT_MN;
            g.C = Xk; g.ldc = T;
            g.krange = KR_A_UPPER;
            GPB_TRY(gemm(s, g));
            GPB_TRY(copy2d(s, nbk, T, Xk, T, B + j0 * ldb, ldb));
            if (j0 > 0) {
                GemmDesc u;  // B[0:j0, :] -= L[j0:j0+nbk, 0:j0]^T * X_k
                u.M = j0; u.N = T; u.K = nbk;
                u.A = L + j0 * lda; u.lda = lda; u.a_layout = LAYOUT_MN;
                u.B = Xk; u.ldb = T; u.b_layout = LAYOUT_MN;
                u.C = B; u.ldc = ldb; u.alpha = -1.0; u.beta = 1.0;
                GPB_TRY(gemm(s, u));
            }
        }
    }
    return GPB_OK;
}

int trtri_into_upper(stream_t s, int64_t N, double* A, int64_t lda, const FactorWs& ws) {
    const int64_t NB = ws.nb;
    const int64_t nblk = nblocks(N, NB);
    double* Wp = ws.panel;  // block column k of W, rows 0 .. j0+nbk, ld NB
    for (int64_t k = 0; k < nblk; ++k) {
        const int64_t j0 = k * NB;
        const int64_t nbk = (N - j0) < NB ? (N - j0) : NB;
        const double* Dk = ws.Dinv + k * NB * NB;
        const double* DTk = ws.DinvT + k * NB * NB;
        const bool tri8 = j0 > 0 && oz_panels_on(ws, j0, nbk);  // scratch planes: the second digit buffer (potrf's double buffer)
        const int64_t ldq = OZ_MAX_SLICES * NB;
        if (j0 > 0) {
            // W[0:j0, k] = -Acc * inv(L_kk)^T
            if (tri8) {
                GPB_TRY(ozaki_slice(s, j0, NB, NB, A + j0, lda, OZ_MAX_SLICES, ws.oz_q2 + NB * ldq, ldq, ws.oz_scale2 + NB, ws.oz_planes));
                GPB_TRY(oz_tri_product(s, ws, j0, ws.oz_q2 + NB * ldq, ws.oz_scale2 + NB, Dk, ws.oz_q2, ws.oz_scale2, Wp, NB, -1.0,
                                       KR_B_LOWER, false));
            } else {
                GemmDesc g;
                g.M = j0; g.N = nbk; g.K = nbk;
                g.A = A + j0; g.lda = lda; g.B = Dk; g.ldb = NB; g.C = Wp; g.ldc = NB;
                g.alpha = -1.0; g.krange = KR_B_LOWER;
                GPB_TRY(gemm(s, g));
            }
            GPB_TRY(copy2d(s, j0, nbk, Wp, NB, A + j0, lda));
        }
        const int64_t right = N - j0 - nbk;
        if (right <= 0) continue;
        GPB_TRY(copy2d(s, nbk, nbk, DTk, NB, Wp + j0 * NB, NB));  // W_kk = inv(L_kk)^T
        const double* Lp = A + (j0 + nbk) * lda + j0;            // L[k+1:, k]
        if (j0 > 0) {
            // Acc[0:j0, k+1:] += W[0:j0,k] * L[k+1:,k]^T   (nbk == NB here: block k is not the last one)
            if (oz_on(ws, j0)) {  // digit planes indexed by GLOBAL row: W rows [0, j0), L rows [j0 + nbk, N)
                const int planes = OZ_MAX_SLICES;
                GPB_TRY(ozaki_slice(s, j0, NB, NB, Wp, NB, planes, ws.oz_q, ldq, ws.oz_scale, ws.oz_planes));
                GPB_TRY(ozaki_slice(s, right, NB, NB, Lp, lda, planes, ws.oz_q + (j0 + nbk) * ldq, ldq, ws.oz_scale + j0 + nbk, ws.oz_planes));
                OzakiGemmDesc u;
                u.M = j0; u.N = right; u.K = NB; u.nslices = planes; u.nslices_dev = ws.oz_planes;
                u.Qa = ws.oz_q; u.ldqa = ldq; u.sa = ws.oz_scale;
                u.Qb = ws.oz_q + (j0 + nbk) * ldq; u.ldqb = ldq; u.sb = ws.oz_scale + j0 + nbk;
                u.C = A + (j0 + nbk); u.ldc = lda; u.alpha = 1.0;
                GPB_TRY(ozaki_gemm(s, u));
            } else {
                GemmDesc u;
                u.M = j0; u.N = right; u.K = nbk;
                u.A = Wp; u.lda = NB; u.B = Lp; u.ldb = lda;
                u.C = A + (j0 + nbk); u.ldc = lda; u.beta = 1.0;
                GPB_TRY(gemm(s, u));
            }
        }
        // Acc[k, k+1:] = W_kk * L[k+1:,k]^T   (first write of that block row)
        if (tri8 && right >= OZ_MIN_ROWS) {  // L's planes are the ones the update above sliced (rows j0 + nbk .. N of the first buffer)
            GPB_TRY(oz_tri_product(s, ws, right, ws.oz_q + (j0 + nbk) * ldq, ws.oz_scale + j0 + nbk, DTk, ws.oz_q2, ws.oz_scale2,
                                   A + j0 * lda + (j0 + nbk), lda, 1.0, KR_A_UPPER, true));
        } else {
            GemmDesc v;
            v.M = nbk; v.N = right; v.K = nbk;
            v.A = Wp + j0 * NB; v.lda = NB; v.B = Lp; v.ldb = lda;
            v.C = A + j0 * lda + (j0 + nbk); v.ldc = lda; v.beta = 0.0;
            v.krange = KR_A_UPPER;
            GPB_TRY(gemm(s, v));
        }
    }
    return GPB_OK;
}

int lauum_upper(stream_t s, int64_t N, double* A, int64_t lda, const FactorWs& ws) {
    const int64_t NB = ws.nb;
    const int64_t nblk = nblocks(N, NB);
    for (int64_t k = 0; k < nblk; ++k) {
        const int64_t j0 = k * NB;
        const int64_t nbk = (N - j0) < NB ? (N - j0) : NB;
        const double* DTk = ws.DinvT + k * NB * NB;
        double* P = A + j0;  // W[0:j0, k], row stride lda
        if (j0 > 0) {
            // S[0:j0, 0:j0] (strictly-upper blocks) += P P^T   and   Sdiag[j] += P_j P_j^T, j < k
            bool diag_done = false;
            if (oz_on(ws, j0)) {  // a ragged last block (nbk < NB) is zero-padded to the next multiple of 128 digits
                const int planes = OZ_MAX_SLICES;
                const int64_t ldq = OZ_MAX_SLICES * NB;
                const int64_t kp = align_up(nbk, 128);
                GPB_TRY(ozaki_slice(s, j0, nbk, kp, P, lda, planes, ws.oz_q, ldq, ws.oz_scale, ws.oz_planes));
                OzakiGemmDesc g;
                g.M = j0; g.N = j0; g.K = kp; g.nslices = planes; g.nslices_dev = ws.oz_planes;
                g.Qa = ws.oz_q; g.ldqa = ldq; g.sa = ws.oz_scale; g.Qb = ws.oz_q; g.ldqb = ldq; g.sb = ws.oz_scale;
                g.C = A; g.ldc = lda; g.alpha = 1.0; g.mask = MASK_BLOCK_STRICT_UPPER; g.mask_nb = NB;
                if (ozaki_supports_extensions()) {  // the same launch also accumulates the diagonal blocks (into ws.Sdiag)
                    g.mask = MASK_BLOCK_UPPER_DIAG_TO_C2; g.C2 = ws.Sdiag;
                    diag_done = true;
                }
                GPB_TRY(ozaki_gemm(s, g));
            } else {
                GemmDesc g;
                g.M = j0; g.N = j0; g.K = nbk;
                g.A = P; g.lda = lda; g.B = P; g.ldb = lda; g.C = A; g.ldc = lda;
                g.beta = 1.0; g.mask = MASK_BLOCK_STRICT_UPPER; g.mask_nb = NB;
                GPB_TRY(gemm(s, g));
            }
            if (!diag_done) {  // diagonal blocks on the DMMA pipe (batched)
                GemmDesc b;
                b.M = NB; b.N = NB; b.K = nbk;
                b.A = P; b.lda = lda; b.B = P; b.ldb = lda; b.C = ws.Sdiag; b.ldc = NB;
                b.beta = 1.0; b.batch = (int)k; b.strideA = NB * lda; b.strideB = NB * lda; b.strideC = NB * NB;
                GPB_TRY(gemm(s, b));
            }
            // S[0:j0, k] = W[0:j0,k] * inv(L_kk)          (B operand = DinvT_k, upper triangular)
            if (oz_panels_on(ws, j0, nbk)) {  // P's planes are still in the first digit buffer (sliced for P P^T above)
                GPB_TRY(oz_tri_product(s, ws, j0, ws.oz_q, ws.oz_scale, DTk, ws.oz_q2, ws.oz_scale2, ws.panel, NB, 1.0, KR_B_UPPER, false));
            } else {
                GemmDesc c;
                c.M = j0; c.N = nbk; c.K = nbk;
                c.A = P; c.lda = lda; c.B = DTk; c.ldb = NB; c.C = ws.panel; c.ldc = NB;
                c.krange = KR_B_UPPER;
                GPB_TRY(gemm(s, c));
            }
            GPB_TRY(copy2d(s, j0, nbk, ws.panel, NB, P, lda));
        }
        // S_kk = inv(L_kk)^T inv(L_kk)
        GemmDesc d;
        d.M = nbk; d.N = nbk; d.K = nbk;
        d.A = DTk; d.lda = NB; d.B = DTk; d.ldb = NB; d.C = ws.Sdiag + k * NB * NB; d.ldc = NB;
        d.krange = KR_A_UPPER;  // inv(L_kk)^T is upper triangular: row m of the left operand starts at column m
        GPB_TRY(gemm(s, d));
    }
    return GPB_OK;
}

// -------------------------------------------------------------------------------------------
// exact-GP objective
// -------------------------------------------------------------------------------------------
int mll_forward(stream_t s, const MllArgs& a, const FactorWs& ws, double* value_out, double* alpha_out, int* info) {
    if (a.N <= 0 || a.D <= 0 || !a.X || !a.y || !a.ell || !a.variance || !a.obs_stddev || !a.Sigma || !value_out ||
        !alpha_out || !info)
        return GPB_ERR_INVALID;
    const int64_t N = a.N, NB = ws.nb;
    double* dvec = ws.vec;                        // d = y - m
    double* wvec = ws.vec + align_up(N, NB);      // w = L^-1 d
    double* half_logdet = ws.scal;
    double* quad = ws.scal + 1;
    GPB_TRY(sub_scalar(s, N, a.y, a.mean_const, dvec));
    GPB_TRY(factor_set_planes(s, ws, N, a.variance, a.obs_stddev, a.jitter));  // conditioning guard of the int8 updates
    GramDesc g;
    g.kind = a.kind; g.N = N; g.M = N; g.D = a.D;
    g.X = a.X; g.ldx = a.ldx; g.Z = a.X; g.ldz = a.ldx;
    g.ell = a.ell; g.ell_is_scalar = a.ell_is_scalar; g.variance = a.variance;
    g.K = a.Sigma; g.ldk = a.lds; g.lower_only = 1;
    g.diag_add = a.jitter; g.diag_add_sq = a.obs_stddev;
    GPB_TRY(gram(s, g));
    GPB_TRY(potrf_lower(s, N, a.Sigma, a.lds, ws, info));
    GPB_TRY(sum_log_diag(s, N, a.Sigma, a.lds, half_logdet));
    GPB_TRY(copy2d(s, 1, N, dvec, N, wvec, N));
    GPB_TRY(trsv_lower(s, N, a.Sigma, a.lds, ws, wvec, 0));
    GPB_TRY(dot(s, N, wvec, wvec, quad));
    GPB_TRY(copy2d(s, 1, N, wvec, N, alpha_out, N));
    GPB_TRY(trsv_lower(s, N, a.Sigma, a.lds, ws, alpha_out, 1));
    GPB_TRY(mll_value(s, N, half_logdet, quad, info, value_out));
    return GPB_OK;
}

int mll_backward(stream_t s, const MllArgs& a, const FactorWs& ws, const double* alpha, const double* gout,
                 double* g_ell, double* g_var, double* g_obs_stddev, double* g_mean) {
    if (a.N <= 0 || !a.Sigma || !alpha || !ws.Sdiag || !ws.partials) return GPB_ERR_INVALID;
    // ws.oz_planes still holds the plane count mll_forward's guard chose for this Sigma ("ws exactly as mll_forward left it")
    GPB_TRY(trtri_into_upper(s, a.N, a.Sigma, a.lds, ws));
    GPB_TRY(lauum_upper(s, a.N, a.Sigma, a.lds, ws));
    MllBwdDesc d;
    d.kind = a.kind; d.N = a.N; d.D = a.D; d.nb = ws.nb;
    d.X = a.X; d.ldx = a.ldx; d.alpha = alpha;
    d.S = a.Sigma; d.lds = a.lds; d.Sdiag = ws.Sdiag;
    d.ell = a.ell; d.ell_is_scalar = a.ell_is_scalar; d.variance = a.variance; d.obs_stddev = a.obs_stddev;
    d.gout = gout; d.partials = ws.partials;
    d.g_ell = g_ell; d.g_var = g_var; d.g_obs_stddev = g_obs_stddev; d.g_mean = g_mean;
    return mll_bwd(s, d);
}

}  // namespace gpb
