// FP64 tensor-core GEMM for sm_100a:  C = beta*C + alpha * A * B^T   (DMMA.8x8x4 via mma.sync m8n8k4)
//
// This single kernel carries every O(N^3) phase of the hot path (Cholesky trailing SYRK and
// panel TRSM-by-inverse, TRTRI, LAUUM, SGPR whitening / SYRK / adjoint GEMM); the reference
// reaches the same arithmetic through jnp.linalg.cholesky / jsp.linalg.solve_triangular /
// jnp.matmul (gpjax/linalg/operations.py:55,107; gpjax/objectives.py:387-404).
//
// Design (B200: 128 FP64 flop/clk/SM, tensor == vector peak, so the kernel is built to keep the
// DMMA pipe issue-saturated while everything else hides behind it):
//   * CTA tile 128x64, 8 warps (4x2), warp tile 32x32 -> 16 DMMA accumulator tiles (64 regs),
//     <=128 regs/thread so TWO CTAs are resident per SM: one CTA's C read-modify-write epilogue
//     overlaps the other's main loop (the rank-256 updates are otherwise ~20 % epilogue).
//   * K is consumed in 16-wide chunks through a 3-stage cp.async (LDGSTS) shared-memory ring;
//     out-of-range rows / K tails are zero-filled by the copy itself (src-size operand).
//   * Shared-memory rows are padded by 4 doubles (32 B) so the 8x4 / 4x8 fragment reads of one
//     half-warp hit all 32 banks exactly once for both operand layouts.
//   * Output masks (triangular / block-triangular) and triangular K-range skipping are applied
//     per tile so SYRK / TRMM-shaped work never touches the dead half.
#include "common.cuh"

namespace gpb {

namespace {

constexpr int BK = 16;
constexpr int STAGES = 3;
constexpr int NTHREADS = 256;
constexpr int PAD = 4;

struct GemmParams {
    int64_t M, N, K;
    const double* A;
    int64_t lda;
    const double* B;
    int64_t ldb;
    double* C;
    int64_t ldc;
    double alpha, beta;
    int mask;
    int64_t mask_row0, mask_col0, mask_nb;
    int krange;
    int64_t kr_off;
    int64_t strideA, strideB, strideC;
    int64_t tiles_n;
    int a_vec16, b_vec16, c_vec16;
};

template <int ROWS, int LAY>
struct TileGeom {
    // LAYOUT_K : smem [ROWS][BK+PAD];  LAYOUT_MN: smem [BK][ROWS+PAD]
    static constexpr int LD = (LAY == LAYOUT_K) ? (BK + PAD) : (ROWS + PAD);
    static constexpr int SIZE = (LAY == LAYOUT_K) ? ROWS * (BK + PAD) : BK * (ROWS + PAD);
};

// Copy one ROWS x BK operand tile into shared memory (zero-filling everything out of range).
template <int ROWS, int LAY>
__device__ __forceinline__ void load_tile(double* __restrict__ sm, const double* __restrict__ G,
                                          int64_t ld, int64_t row0, int64_t nrows, int64_t k0,
                                          int64_t kend, bool vec16, int tid) {
    constexpr int LD = TileGeom<ROWS, LAY>::LD;
    if (LAY == LAYOUT_K) {
        constexpr int CHUNKS = ROWS * (BK / 2);
#pragma unroll
        for (int id = tid; id < CHUNKS; id += NTHREADS) {
            int r = id / (BK / 2), ch = id % (BK / 2);
            int64_t gm = row0 + r;
            int64_t k = k0 + ch * 2;
            int64_t left = kend - k;
            int nb = (gm < nrows) ? (left >= 2 ? 16 : (left == 1 ? 8 : 0)) : 0;
            const double* src = (nb > 0) ? (G + gm * ld + k) : G;
            double* dst = sm + r * LD + ch * 2;
            if (vec16) {
                cp_async16(dst, src, nb);
            } else {
                cp_async8(dst, src, nb >= 8 ? 8 : 0);
                cp_async8(dst + 1, (nb == 16) ? (src + 1) : G, nb == 16 ? 8 : 0);
            }
        }
    } else {
        constexpr int CPR = ROWS / 2;  // 16-byte chunks per k-row
        constexpr int CHUNKS = BK * CPR;
#pragma unroll
        for (int id = tid; id < CHUNKS; id += NTHREADS) {
            int kr = id / CPR, ch = id % CPR;
            int64_t k = k0 + kr;
            int64_t m = row0 + ch * 2;
            int64_t left = nrows - m;
            int nb = (k < kend) ? (left >= 2 ? 16 : (left == 1 ? 8 : 0)) : 0;
            const double* src = (nb > 0) ? (G + k * ld + m) : G;
            double* dst = sm + kr * LD + ch * 2;
            if (vec16) {
                cp_async16(dst, src, nb);
            } else {
                cp_async8(dst, src, nb >= 8 ? 8 : 0);
                cp_async8(dst + 1, (nb == 16) ? (src + 1) : G, nb == 16 ? 8 : 0);
            }
        }
    }
}

__device__ __forceinline__ bool mask_keep(int mask, int64_t r, int64_t c, int64_t nb) {
    switch (mask) {
        case MASK_LOWER: return r >= c;
        case MASK_UPPER: return r <= c;
        case MASK_BLOCK_STRICT_UPPER: return (r / nb) < (c / nb);
        case MASK_BLOCK_STRICT_LOWER: return (r / nb) > (c / nb);
        default: return true;
    }
}

template <int BM, int BN, int WM, int WN, int ALAY, int BLAY>
__global__ void __launch_bounds__(NTHREADS, 2) gemm_f64_kernel(const GemmParams p) {
    constexpr int WARPS_N = BN / WN;
    constexpr int TM = WM / 8, TN = WN / 8;
    using GA = TileGeom<BM, ALAY>;
    using GB = TileGeom<BN, BLAY>;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * GA::SIZE;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / WARPS_N) * WM;
    const int wn0 = (warp % WARPS_N) * WN;

    const int64_t tile_m = blockIdx.x / p.tiles_n;
    const int64_t tile_n = blockIdx.x % p.tiles_n;
    const int64_t m0 = tile_m * BM, n0 = tile_n * BN;
    const int64_t mend = min(m0 + (int64_t)BM, p.M), nend = min(n0 + (int64_t)BN, p.N);

    // ---- tile-level mask: skip tiles with no live element -------------------------------
    if (p.mask != MASK_NONE) {
        int64_t rmin = p.mask_row0 + m0, rmax = p.mask_row0 + mend - 1;
        int64_t cmin = p.mask_col0 + n0, cmax = p.mask_col0 + nend - 1;
        bool live = true;
        if (p.mask == MASK_LOWER) live = rmax >= cmin;
        else if (p.mask == MASK_UPPER) live = rmin <= cmax;
        else if (p.mask == MASK_BLOCK_STRICT_UPPER) live = (rmin / p.mask_nb) < (cmax / p.mask_nb);
        else if (p.mask == MASK_BLOCK_STRICT_LOWER) live = (rmax / p.mask_nb) > (cmin / p.mask_nb);
        if (!live) return;
    }

    const double* A = p.A + (int64_t)blockIdx.y * p.strideA;
    const double* B = p.B + (int64_t)blockIdx.y * p.strideB;
    double* C = p.C + (int64_t)blockIdx.y * p.strideC;

    // ---- K range (triangular operands: structural zeros must be physically zero) ---------
    int64_t kbeg = 0, kend = p.K;
    if (p.krange == KR_B_LOWER) kend = min(p.K, nend + p.kr_off);
    else if (p.krange == KR_B_UPPER) kbeg = max((int64_t)0, n0 + p.kr_off);
    else if (p.krange == KR_A_LOWER) kend = min(p.K, mend + p.kr_off);
    else if (p.krange == KR_A_UPPER) kbeg = max((int64_t)0, m0 + p.kr_off);
    kbeg = (kbeg / BK) * BK;
    if (kend < kbeg) kend = kbeg;
    const int nk = (int)((kend - kbeg + BK - 1) / BK);

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool av = p.a_vec16 != 0, bv = p.b_vec16 != 0;

    // ---- prologue ------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            load_tile<BM, ALAY>(As + s * GA::SIZE, A, p.lda, m0, p.M, kbeg + (int64_t)s * BK, kend, av, tid);
            load_tile<BN, BLAY>(Bs + s * GB::SIZE, B, p.ldb, n0, p.N, kbeg + (int64_t)s * BK, kend, bv, tid);
        }
        cp_async_commit();
    }

    const int fr = lane >> 2, fc = lane & 3;

    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = it + STAGES - 1;
            if (nx < nk) {
                int st = nx % STAGES;
                load_tile<BM, ALAY>(As + st * GA::SIZE, A, p.lda, m0, p.M, kbeg + (int64_t)nx * BK, kend, av, tid);
                load_tile<BN, BLAY>(Bs + st * GB::SIZE, B, p.ldb, n0, p.N, kbeg + (int64_t)nx * BK, kend, bv, tid);
            }
            cp_async_commit();
        }
        const double* as = As + (it % STAGES) * GA::SIZE;
        const double* bs = Bs + (it % STAGES) * GB::SIZE;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                if (ALAY == LAYOUT_K) af[i] = as[(wm0 + i * 8 + fr) * GA::LD + kk * 4 + fc];
                else af[i] = as[(kk * 4 + fc) * GA::LD + wm0 + i * 8 + fr];
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                if (BLAY == LAYOUT_K) bf[j] = bs[(wn0 + j * 8 + fr) * GB::LD + kk * 4 + fc];
                else bf[j] = bs[(kk * 4 + fc) * GB::LD + wn0 + j * 8 + fr];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ------------------------------------------------------------------------
    const double alpha = p.alpha, beta = p.beta;
    const bool cvec = p.c_vec16 != 0;
    const bool use_mask = p.mask != MASK_NONE;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t row = m0 + wm0 + i * 8 + fr;
        if (row >= p.M) continue;
        double* crow = C + row * p.ldc;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int64_t col = n0 + wn0 + j * 8 + fc * 2;
            if (col >= p.N) continue;
            bool ok0 = true, ok1 = (col + 1 < p.N);
            if (use_mask) {
                ok0 = mask_keep(p.mask, p.mask_row0 + row, p.mask_col0 + col, p.mask_nb);
                ok1 = ok1 && mask_keep(p.mask, p.mask_row0 + row, p.mask_col0 + col + 1, p.mask_nb);
            }
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
            if (ok0 && ok1 && cvec) {
                double2* ptr = reinterpret_cast<double2*>(crow + col);
                if (beta != 0.0) {
                    double2 old = *ptr;
                    v0 += beta * old.x;
                    v1 += beta * old.y;
                }
                *ptr = make_double2(v0, v1);
            } else {
                if (ok0) {
                    if (beta != 0.0) v0 += beta * crow[col];
                    crow[col] = v0;
                }
                if (ok1) {
                    if (beta != 0.0) v1 += beta * crow[col + 1];
                    crow[col + 1] = v1;
                }
            }
        }
    }
}

template <int BM, int BN, int WM, int WN, int ALAY, int BLAY>
int launch_gemm(cudaStream_t st, const GemmParams& p, int batch) {
    using GA = TileGeom<BM, ALAY>;
    using GB = TileGeom<BN, BLAY>;
    constexpr size_t smem = sizeof(double) * STAGES * (GA::SIZE + GB::SIZE);
    auto kern = gemm_f64_kernel<BM, BN, WM, WN, ALAY, BLAY>;
    static bool configured = false;  // per-instantiation, idempotent
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return GPB_ERR_LAUNCH;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured = true;
    }
    int64_t tiles_m = (p.M + BM - 1) / BM;
    int64_t ntiles = tiles_m * p.tiles_n;
    if (ntiles <= 0 || batch <= 0) return GPB_OK;
    if (ntiles > 2147483647LL || batch > 65535) return GPB_ERR_UNSUPPORTED;
    dim3 grid((unsigned)ntiles, (unsigned)batch, 1);
    kern<<<grid, NTHREADS, smem, st>>>(p);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace

int gemm(stream_t s, const GemmDesc& d) {
    if (d.M < 0 || d.N < 0 || d.K < 0) return GPB_ERR_INVALID;
    if (d.M == 0 || d.N == 0 || d.batch == 0) return GPB_OK;
    if (!d.C || (d.K > 0 && (!d.A || !d.B))) return GPB_ERR_INVALID;
    constexpr int BM = 128, BN = 64, WM = 32, WN = 32;
    GemmParams p;
    p.M = d.M; p.N = d.N; p.K = d.K;
    p.A = d.A; p.lda = d.lda; p.B = d.B; p.ldb = d.ldb; p.C = d.C; p.ldc = d.ldc;
    p.alpha = d.alpha; p.beta = d.beta;
    p.mask = d.mask; p.mask_row0 = d.mask_row0; p.mask_col0 = d.mask_col0;
    p.mask_nb = d.mask_nb > 0 ? d.mask_nb : 1;
    p.krange = d.krange; p.kr_off = d.kr_off;
    p.strideA = d.strideA; p.strideB = d.strideB; p.strideC = d.strideC;
    p.tiles_n = (d.N + BN - 1) / BN;
    auto al16 = [](const void* ptr, int64_t ld, int64_t stride) {
        return ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld % 2 == 0) && (stride % 2 == 0);
    };
    p.a_vec16 = al16(d.A, d.lda, d.strideA);
    p.b_vec16 = al16(d.B, d.ldb, d.strideB);
    p.c_vec16 = al16(d.C, d.ldc, d.strideC);
    cudaStream_t st = to_stream(s);
    if (d.a_layout == LAYOUT_K && d.b_layout == LAYOUT_K)
        return launch_gemm<BM, BN, WM, WN, LAYOUT_K, LAYOUT_K>(st, p, d.batch);
    if (d.a_layout == LAYOUT_K && d.b_layout == LAYOUT_MN)
        return launch_gemm<BM, BN, WM, WN, LAYOUT_K, LAYOUT_MN>(st, p, d.batch);
    if (d.a_layout == LAYOUT_MN && d.b_layout == LAYOUT_K)
        return launch_gemm<BM, BN, WM, WN, LAYOUT_MN, LAYOUT_K>(st, p, d.batch);
    if (d.a_layout == LAYOUT_MN && d.b_layout == LAYOUT_MN)
        return launch_gemm<BM, BN, WM, WN, LAYOUT_MN, LAYOUT_MN>(st, p, d.batch);
    return GPB_ERR_INVALID;
}

}  // namespace gpb
