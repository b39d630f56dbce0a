// FP64 tensor-core GEMM for sm_100a:  C = beta*C + alpha * A * B^T   (DMMA.8x8x4 via mma.sync m8n8k4)
//
// This single kernel carries every O(N^3) phase of the hot path (Cholesky trailing SYRK and
// panel TRSM-by-inverse, TRTRI, LAUUM, SGPR whitening / SYRK / adjoint GEMM); the reference
// reaches the same arithmetic through jnp.linalg.cholesky / jsp.linalg.solve_triangular /
// jnp.matmul (gpjax/linalg/operations.py:55,107; gpjax/objectives.py:387-404).
//
// Design (B200: 128 FP64 flop/clk/SM, tensor == vector peak = 37.2 TF/s, so the kernel is built to keep
// the DMMA pipe issue-saturated while everything else hides behind it; all choices below were measured
// with scripts/gemm_bench.py + ncu, see profiles/r01_summary.md):
//   * CTA tile 128x64, 4 warps (4x1), warp tile 32x64 -> 32 DMMA accumulator tiles per warp (12 fragment
//     loads per 32 DMMAs), ~230 regs, TWO CTAs resident per SM with independent barriers: one CTA's C
//     read-modify-write epilogue and chunk barriers overlap the other's main loop.  (8 warps x 32x32:
//     -4 %; a 128x128 CTA with one CTA/SM: -8 %.)  Launches of a few tiles only (<= 37 such CTAs) take the same kernel
//     instantiated with quarter tiles (64x32, two warps of 32x32): see gemm() at the end of the file.
//   * K is consumed in 16-wide chunks through a 3-stage cp.async (LDGSTS) shared-memory ring; interior
//     tiles use a precomputed-pointer fast path (3 instructions per 16-byte copy), edge tiles a fully
//     predicated path whose zero-fill comes from the copy's src-size operand.
//   * Shared-memory rows are padded by 4 doubles (32 B) so the 8x4 / 4x8 fragment reads of one
//     half-warp hit all 32 banks exactly once for both operand layouts; fragments for k-slice kk+1 are
//     fetched while slice kk feeds the tensor pipe.
//   * Epilogue: the C tile is prefetched to L2 at kernel start (beta != 0) and interior tiles load all
//     C fragments of a half tile back to back (one memory round trip per half instead of one per
//     fragment: rank-256 updates went from 75 % to 89 % DMMA-pipe utilisation).
//   * Grid: tile rows are folded (y, T-1-y) so triangular / block-triangular masks launch no dead CTAs,
//     and live columns are walked in chunks of 64 tiles (blockIdx.z) so the B rows of a chunk stay
//     L2-resident while all tile rows sweep them (DRAM reads of a 44.5k^2 rank-512 update: 41 GB -> 10 GB).
//   * Triangular K-range skipping for operands with physically-zero triangles (explicit inverses).
#include <cstdlib>
#include "common.cuh"

namespace gpb {

namespace {

constexpr int BK = 16;
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
    int64_t tiles_m, tiles_n;
    int a_vec16, b_vec16, c_vec16;
    int batch;
};

// Shared-memory tile geometry.  Two conflict-free schemes for the 8x4 / 4x8 fragment reads:
//   SWZ = false: rows padded by 4 doubles (32 B)          -> 30.7 KB per 128x64 stage
//   SWZ = true : no padding, the 32-byte segments of a row are XOR-permuted with (row & 3) -> 24.6 KB per stage,
//                which is what lets THREE CTAs share an SM (3 x 73.7 KB).
template <int ROWS, int LAY, bool SWZ>
struct TileGeom {
    // LAYOUT_K : smem [ROWS][BK(+PAD)];  LAYOUT_MN: smem [BK][ROWS(+PAD)]
    static constexpr int P = SWZ ? 0 : PAD;
    static constexpr int LD = (LAY == LAYOUT_K) ? (BK + P) : (ROWS + P);
    static constexpr int SIZE = (LAY == LAYOUT_K) ? ROWS * (BK + P) : BK * (ROWS + P);
    // element (major, minor): LAYOUT_K -> (row, k), LAYOUT_MN -> (k, row)
    __device__ __forceinline__ static int index(int major, int minor) {
        if (SWZ) return major * LD + ((((minor >> 2) ^ (major & 3)) << 2) | (minor & 3));
        return major * LD + minor;
    }
};

// Copy one ROWS x BK operand tile into shared memory (zero-filling everything out of range).
template <int ROWS, int LAY, int NT, bool SWZ>
__device__ __forceinline__ void load_tile(double* __restrict__ sm, const double* __restrict__ G,
                                          int64_t ld, int64_t row0, int64_t nrows, int64_t k0,
                                          int64_t kend, bool vec16, int tid) {
    using G_ = TileGeom<ROWS, LAY, SWZ>;
    if (LAY == LAYOUT_K) {
        constexpr int CHUNKS = ROWS * (BK / 2);
#pragma unroll
        for (int id = tid; id < CHUNKS; id += NT) {
            int r = id / (BK / 2), ch = id % (BK / 2);
            int64_t gm = row0 + r;
            int64_t k = k0 + ch * 2;
            int64_t left = kend - k;
            int nb = (gm < nrows) ? (left >= 2 ? 16 : (left == 1 ? 8 : 0)) : 0;
            const double* src = (nb > 0) ? (G + gm * ld + k) : G;
            double* dst = sm + G_::index(r, ch * 2);
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
        for (int id = tid; id < CHUNKS; id += NT) {
            int kr = id / CPR, ch = id % CPR;
            int64_t k = k0 + kr;
            int64_t m = row0 + ch * 2;
            int64_t left = nrows - m;
            int nb = (k < kend) ? (left >= 2 ? 16 : (left == 1 ? 8 : 0)) : 0;
            const double* src = (nb > 0) ? (G + k * ld + m) : G;
            double* dst = sm + G_::index(kr, ch * 2);
            if (vec16) {
                cp_async16(dst, src, nb);
            } else {
                cp_async8(dst, src, nb >= 8 ? 8 : 0);
                cp_async8(dst + 1, (nb == 16) ? (src + 1) : G, nb == 16 ? 8 : 0);
            }
        }
    }
}


// Fast-path copy of one operand tile: every row of the tile is inside the matrix, the whole BK-wide
// K chunk is inside [kbeg, kend) and rows are 16-byte aligned, so each thread issues its 16-byte
// cp.async's from a precomputed base pointer with compile-time strides (3 instructions per copy
// instead of ~20 for the fully predicated path).
template <int ROWS, int LAY, int NT, bool SWZ>
struct FastPlan {
    using G_ = TileGeom<ROWS, LAY, SWZ>;
    static constexpr int NCOPY = ROWS * (BK / 2) / NT;
    static_assert(ROWS * (BK / 2) % NT == 0, "tile copies must divide evenly over the threads");
    static constexpr int CPM = (LAY == LAYOUT_K) ? (BK / 2) : (ROWS / 2);  // 16-byte chunks per major index
    static constexpr int MSTEP = NT / CPM;                                 // major-index step between copies
    const double* src;   // global address of this thread's first chunk at k = 0
    int64_t src_step;    // elements between consecutive chunks of this thread
    int64_t k_step;      // elements per unit of k
    int major0, minor0;  // (row, k) resp. (k, row) of the first chunk inside the tile
    __device__ __forceinline__ void init(const double* G, int64_t ld, int64_t row0, int tid) {
        major0 = tid / CPM;
        minor0 = (tid % CPM) * 2;
        if (LAY == LAYOUT_K) {
            src = G + (row0 + major0) * ld + minor0;
            k_step = 1;
        } else {
            src = G + (int64_t)major0 * ld + row0 + minor0;
            k_step = ld;
        }
        src_step = (int64_t)MSTEP * ld;
    }
    __device__ __forceinline__ void issue(double* stage, int64_t k0) const {
        const double* s0 = src + k0 * k_step;
#pragma unroll
        for (int i = 0; i < NCOPY; ++i)
            cp_async16(stage + G_::index(major0 + i * MSTEP, minor0), s0 + i * src_step, 16);
    }
};

__device__ __forceinline__ bool mask_keep(int mask, int64_t r, int64_t c, int64_t nb) {
    switch (mask) {
        case MASK_LOWER: return r >= c;
        case MASK_UPPER: return r <= c;
        case MASK_BLOCK_STRICT_UPPER: return (r / nb) < (c / nb);
        case MASK_BLOCK_STRICT_LOWER: return (r / nb) > (c / nb);
        default: return true;
    }
}


__host__ __device__ inline int64_t floor_div(int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return -floor_div(-a, b); }

// Column-tile range [lo, hi) of tile-row `tm` that can contain live (unmasked) elements.  May be a
// slight over-estimate at the ragged last column tile; the kernel re-checks liveness exactly.
template <int BM, int BN>
__host__ __device__ inline void live_range(const GemmParams& p, int64_t tm, int64_t& lo, int64_t& hi) {
    lo = 0;
    hi = p.tiles_n;
    if (p.mask == MASK_NONE) return;
    const int64_t rend = (tm + 1) * BM < p.M ? (tm + 1) * BM : p.M;
    const int64_t rmin = p.mask_row0 + tm * BM, rmax = p.mask_row0 + rend - 1;
    int64_t t;
    switch (p.mask) {
        case MASK_LOWER:  // live iff rmax >= mask_col0 + tn*BN
            t = floor_div(rmax - p.mask_col0, BN) + 1;
            hi = t < 0 ? 0 : (t < hi ? t : hi);
            break;
        case MASK_UPPER:  // live iff rmin <= mask_col0 + (tn+1)*BN - 1
            t = ceil_div(rmin - p.mask_col0 + 1, BN) - 1;
            lo = t < 0 ? 0 : (t < hi ? t : hi);
            break;
        case MASK_BLOCK_STRICT_UPPER:  // live iff (rmin/nb + 1)*nb <= mask_col0 + (tn+1)*BN - 1
            t = ceil_div((rmin / p.mask_nb + 1) * p.mask_nb - p.mask_col0 + 1, BN) - 1;
            lo = t < 0 ? 0 : (t < hi ? t : hi);
            break;
        case MASK_BLOCK_STRICT_LOWER:  // live iff mask_col0 + tn*BN < (rmax/nb)*nb
            t = ceil_div((rmax / p.mask_nb) * p.mask_nb - p.mask_col0, BN);
            hi = t < 0 ? 0 : (t < hi ? t : hi);
            break;
        default: break;
    }
}

template <int BM, int BN, int WM, int WN, int STAGES, int MINB, bool SWZ, bool BULK, int ALAY, int BLAY>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB) gemm_f64_kernel(const GemmParams p) {
    static_assert(!(BULK && SWZ), "bulk row copies need the contiguous (padded) row layout");
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int WARPS_N = BN / WN;
    constexpr int TM = WM / 8, TN = WN / 8;
    using GA = TileGeom<BM, ALAY, SWZ>;
    using GB = TileGeom<BN, BLAY, SWZ>;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * GA::SIZE;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (GA::SIZE + GB::SIZE));  // [STAGES] (BULK only)

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp / WARPS_N) * WM;
    const int wn0 = (warp % WARPS_N) * WN;
    if (BULK) {
        if (tid == 0) {
#pragma unroll
            for (int st = 0; st < STAGES; ++st) mbar_init(full + st, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }

    // Tile enumeration without dead CTAs: tile-row y is folded with tile-row T-1-y (their live
    // column counts add up to ~const for triangular masks), blockIdx.x walks both live ranges.
    int64_t tile_m = blockIdx.y, tile_n;
    {
        // blockIdx.z = chunk * batch + b: the live columns of a folded row pair are walked in chunks of
        // gridDim.x tiles so that the B rows a chunk touches (gridDim.x * BN * K doubles) stay L2-resident
        // while every tile row sweeps over them (rank-k updates with N >> L2 otherwise re-stream B per row).
        int64_t lo, hi, x = blockIdx.x + (int64_t)gridDim.x * (blockIdx.z / p.batch);
        live_range<BM, BN>(p, tile_m, lo, hi);
        if (x < hi - lo) {
            tile_n = lo + x;
        } else {
            x -= hi - lo;
            const int64_t t2 = p.tiles_m - 1 - tile_m;
            if (t2 == tile_m) return;
            tile_m = t2;
            live_range<BM, BN>(p, tile_m, lo, hi);
            if (x >= hi - lo) return;
            tile_n = lo + x;
        }
    }
    const int64_t m0 = tile_m * BM, n0 = tile_n * BN;
    const int64_t mend = min(m0 + (int64_t)BM, p.M), nend = min(n0 + (int64_t)BN, p.N);

    // ---- tile-level mask: skip dead tiles, detect fully-live ("interior") tiles ---------------
    bool interior = (m0 + BM <= p.M) && (n0 + BN <= p.N) && (p.c_vec16 != 0);
    if (p.mask != MASK_NONE) {
        int64_t rmin = p.mask_row0 + m0, rmax = p.mask_row0 + mend - 1;
        int64_t cmin = p.mask_col0 + n0, cmax = p.mask_col0 + nend - 1;
        bool live = true, full = true;
        if (p.mask == MASK_LOWER) { live = rmax >= cmin; full = rmin >= cmax; }
        else if (p.mask == MASK_UPPER) { live = rmin <= cmax; full = rmax <= cmin; }
        else if (p.mask == MASK_BLOCK_STRICT_UPPER) {
            live = (rmin / p.mask_nb) < (cmax / p.mask_nb);
            full = (rmax / p.mask_nb) < (cmin / p.mask_nb);
        } else if (p.mask == MASK_BLOCK_STRICT_LOWER) {
            live = (rmax / p.mask_nb) > (cmin / p.mask_nb);
            full = (rmin / p.mask_nb) > (cmax / p.mask_nb);
        }
        if (!live) return;
        interior = interior && full;
    }

    const double* A = p.A + (int64_t)(blockIdx.z % p.batch) * p.strideA;
    const double* B = p.B + (int64_t)(blockIdx.z % p.batch) * p.strideB;
    double* C = p.C + (int64_t)(blockIdx.z % p.batch) * p.strideC;
    if (p.beta != 0.0) {
        // pull the C tile towards L2 while the main loop runs: 128 rows x 512 B = 4 lines per row
        for (int id = tid; id < BM * (BN / 16); id += NT) {
            int r = id / (BN / 16), q = id % (BN / 16);
            if (m0 + r < p.M && n0 + q * 16 < p.N)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (m0 + r) * p.ldc + n0 + q * 16));
        }
    }

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
    // CTA-uniform fast-path conditions (see FastPlan)
    const bool a_fast = av && (m0 + BM <= p.M);
    const bool b_fast = bv && (n0 + BN <= p.N);
    FastPlan<BM, ALAY, NT, SWZ> pa;
    FastPlan<BN, BLAY, NT, SWZ> pb;
    pa.init(A, p.lda, m0, tid);
    pb.init(B, p.ldb, n0, tid);

    // Bulk path (CTA-uniform): every chunk of this tile is a full, aligned BK-wide slab, so each operand row is
    // ONE cp.async.bulk (TMA engine) of BK*8 (K layout) or ROWS*8 (MN layout) bytes tracked by the stage's
    // mbarrier; 1-2 copy instructions per thread per chunk instead of 12 LDGSTS.
    const bool use_bulk = BULK && a_fast && b_fast && ((kend - kbeg) % BK == 0);
    auto bulk_stage = [&](int st, int64_t k0) {
        if (tid == 0) mbar_arrive_expect_tx(full + st, (unsigned)((BM + BN) * BK * sizeof(double)));
        double* as_ = As + st * GA::SIZE;
        double* bs_ = Bs + st * GB::SIZE;
        if (ALAY == LAYOUT_K) {
            for (int r = tid; r < BM; r += NT) bulk_g2s(as_ + r * GA::LD, A + (m0 + r) * p.lda + k0, BK * 8, full + st);
        } else {
            for (int kr = tid; kr < BK; kr += NT) bulk_g2s(as_ + kr * GA::LD, A + (k0 + kr) * p.lda + m0, BM * 8, full + st);
        }
        if (BLAY == LAYOUT_K) {
            for (int r = tid; r < BN; r += NT) bulk_g2s(bs_ + r * GB::LD, B + (n0 + r) * p.ldb + k0, BK * 8, full + st);
        } else {
            // spread over the upper lanes so the A and B row copies of an MN/MN tile come from different threads
            for (int kr = NT - 1 - tid; kr < BK; kr += NT) bulk_g2s(bs_ + kr * GB::LD, B + (k0 + kr) * p.ldb + n0, BN * 8, full + st);
        }
    };
    auto load_stage = [&](int st, int64_t k0) {
        if (use_bulk) { bulk_stage(st, k0); return; }
        const bool kfull = (k0 + BK <= kend);
        if (a_fast && kfull) pa.issue(As + st * GA::SIZE, k0);
        else load_tile<BM, ALAY, NT, SWZ>(As + st * GA::SIZE, A, p.lda, m0, p.M, k0, kend, av, tid);
        if (b_fast && kfull) pb.issue(Bs + st * GB::SIZE, k0);
        else load_tile<BN, BLAY, NT, SWZ>(Bs + st * GB::SIZE, B, p.ldb, n0, p.N, k0, kend, bv, tid);
    };

    // ---- prologue ------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, kbeg + (int64_t)s * BK);
        cp_async_commit();
    }

    const int fr = lane >> 2, fc = lane & 3;
    // Per-thread fragment offsets inside a stage: element (row, k) of A / (col, k) of B for
    // row = w?0 + i*8 + fr, k = kk*4 + fc.  Split into an i/j-dependent and a kk-dependent part so the
    // unrolled loop only adds compile-time constants (padded) or a precomputed XOR term (swizzled).
    int a_i[TM], b_j[TN], a_k[BK / 4], b_k[BK / 4];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = wm0 + i * 8 + fr;
        if (ALAY == LAYOUT_K) a_i[i] = row * GA::LD + fc;
        else a_i[i] = SWZ ? ((((row >> 2) ^ fc) << 2) | (row & 3)) : row;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int col = wn0 + j * 8 + fr;
        if (BLAY == LAYOUT_K) b_j[j] = col * GB::LD + fc;
        else b_j[j] = SWZ ? ((((col >> 2) ^ fc) << 2) | (col & 3)) : col;
    }
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
        // LAYOUT_K: k segment kk, permuted with (row & 3) = (fr & 3) when swizzled; LAYOUT_MN: k-row (kk*4 + fc)
        a_k[kk] = (ALAY == LAYOUT_K) ? ((SWZ ? (kk ^ (fr & 3)) : kk) << 2) : (kk * 4 + fc) * GA::LD;
        b_k[kk] = (BLAY == LAYOUT_K) ? ((SWZ ? (kk ^ (fr & 3)) : kk) << 2) : (kk * 4 + fc) * GB::LD;
    }

    for (int it = 0; it < nk; ++it) {
        if (use_bulk) mbar_wait(full + (it % STAGES), (unsigned)((it / STAGES) & 1));
        else cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = it + STAGES - 1;
            if (nx < nk) load_stage(nx % STAGES, kbeg + (int64_t)nx * BK);
            cp_async_commit();
        }
        const double* as = As + (it % STAGES) * GA::SIZE;
        const double* bs = Bs + (it % STAGES) * GB::SIZE;
        double af[2][TM], bf[2][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) af[0][i] = as[a_i[i] + a_k[0]];
#pragma unroll
        for (int j = 0; j < TN; ++j) bf[0][j] = bs[b_j[j] + b_k[0]];
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const int cur = kk & 1, nxt = cur ^ 1;
            if (kk + 1 < BK / 4) {  // fetch the next k4 slice while this one feeds the tensor pipe
#pragma unroll
                for (int i = 0; i < TM; ++i) af[nxt][i] = as[a_i[i] + a_k[kk + 1]];
#pragma unroll
                for (int j = 0; j < TN; ++j) bf[nxt][j] = bs[b_j[j] + b_k[kk + 1]];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ------------------------------------------------------------------------
    const double alpha = p.alpha, beta = p.beta;
    if (interior) {
        // Fast path (tile fully inside the matrix and the mask, 16-byte aligned rows): all C loads
        // of a half-tile are issued back to back (one memory round trip per half instead of one per
        // 8x8 fragment), then blended and stored as 128-bit words.
        double* cbase = C + (m0 + wm0 + fr) * p.ldc + (n0 + wn0 + fc * 2);
        if (beta != 0.0) {
#pragma unroll
            for (int h = 0; h < TM; h += 2) {
                double2 old[2][TN];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j)
                        old[i][j] = *reinterpret_cast<const double2*>(cbase + (int64_t)(h + i) * 8 * p.ldc + j * 8);
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        double2 v;
                        v.x = fma(beta, old[i][j].x, alpha * acc[h + i][j][0]);
                        v.y = fma(beta, old[i][j].y, alpha * acc[h + i][j][1]);
                        *reinterpret_cast<double2*>(cbase + (int64_t)(h + i) * 8 * p.ldc + j * 8) = v;
                    }
            }
        } else {
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    *reinterpret_cast<double2*>(cbase + (int64_t)i * 8 * p.ldc + j * 8) =
                        make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
        }
        return;
    }
    // Generic path: bounds, element masks, unaligned rows.
    const bool use_mask = p.mask != MASK_NONE;
    const bool cvec = p.c_vec16 != 0;
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
                    v0 = fma(beta, old.x, v0);
                    v1 = fma(beta, old.y, v1);
                }
                *ptr = make_double2(v0, v1);
            } else {
                if (ok0) {
                    if (beta != 0.0) v0 = fma(beta, crow[col], v0);
                    crow[col] = v0;
                }
                if (ok1) {
                    if (beta != 0.0) v1 = fma(beta, crow[col + 1], v1);
                    crow[col + 1] = v1;
                }
            }
        }
    }
}

template <int BM, int BN, int WM, int WN, int STAGES, int MINB, bool SWZ, bool BULK, int ALAY, int BLAY>
int launch_gemm(cudaStream_t st, const GemmParams& p, int batch) {
    using GA = TileGeom<BM, ALAY, SWZ>;
    using GB = TileGeom<BN, BLAY, SWZ>;
    constexpr size_t smem = sizeof(double) * STAGES * (GA::SIZE + GB::SIZE) + (BULK ? STAGES * sizeof(uint64_t) : 0);
    auto kern = gemm_f64_kernel<BM, BN, WM, WN, STAGES, MINB, SWZ, BULK, ALAY, BLAY>;
    static PerDeviceOnce configured;  // per instantiation AND per device
    const int dev = current_device();
    if (configured.needed(dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return GPB_ERR_LAUNCH;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured.mark(dev);
    }
    const int64_t T = p.tiles_m;
    if (T <= 0 || p.tiles_n <= 0 || batch <= 0) return GPB_OK;
    const int64_t gy = (T + 1) / 2;
    int64_t gx = 0;
    for (int64_t y = 0; y < gy; ++y) {
        int64_t lo, hi;
        live_range<BM, BN>(p, y, lo, hi);
        int64_t c = hi - lo;
        if (T - 1 - y != y) {
            live_range<BM, BN>(p, T - 1 - y, lo, hi);
            c += hi - lo;
        }
        if (c > gx) gx = c;
    }
    if (gx == 0) return GPB_OK;
    constexpr int64_t CHUNK = 64;  // column tiles per rasterisation chunk (64 * BN rows of B, K doubles each)
    const int64_t nchunks = (gx + CHUNK - 1) / CHUNK;
    const int64_t gxc = nchunks > 1 ? CHUNK : gx;
    if (gy > 65535 || nchunks * batch > 65535) return GPB_ERR_UNSUPPORTED;
    dim3 grid((unsigned)gxc, (unsigned)gy, (unsigned)(nchunks * batch));
    const bool prof = profile_enabled();
    if (prof) profile_gemm_begin(st);
    kern<<<grid, (BM / WM) * (BN / WN) * 32, smem, st>>>(p);
    if (prof) profile_gemm_end(st);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace

static int g_variant = 2;
static const bool g_small_tiles = [] {
    const char* e = std::getenv("GPB_GEMM_SMALL_TILES");
    return !(e && e[0] == '0');
}();
void debug_set_gemm_variant(int v) { g_variant = v; }

template <int BM, int BN, int WM, int WN, int STAGES, int MINB, bool SWZ, bool BULK>
static int dispatch_layouts(cudaStream_t st, GemmParams p, const GemmDesc& d) {
    p.tiles_m = (d.M + BM - 1) / BM;
    p.tiles_n = (d.N + BN - 1) / BN;
    if (d.a_layout == LAYOUT_K && d.b_layout == LAYOUT_K)
        return launch_gemm<BM, BN, WM, WN, STAGES, MINB, SWZ, BULK, LAYOUT_K, LAYOUT_K>(st, p, d.batch);
    if (d.a_layout == LAYOUT_K && d.b_layout == LAYOUT_MN)
        return launch_gemm<BM, BN, WM, WN, STAGES, MINB, SWZ, BULK, LAYOUT_K, LAYOUT_MN>(st, p, d.batch);
    if (d.a_layout == LAYOUT_MN && d.b_layout == LAYOUT_K)
        return launch_gemm<BM, BN, WM, WN, STAGES, MINB, SWZ, BULK, LAYOUT_MN, LAYOUT_K>(st, p, d.batch);
    if (d.a_layout == LAYOUT_MN && d.b_layout == LAYOUT_MN)
        return launch_gemm<BM, BN, WM, WN, STAGES, MINB, SWZ, BULK, LAYOUT_MN, LAYOUT_MN>(st, p, d.batch);
    return GPB_ERR_INVALID;
}

int gemm(stream_t s, const GemmDesc& d) {
    if (d.M < 0 || d.N < 0 || d.K < 0) return GPB_ERR_INVALID;
    if (d.M == 0 || d.N == 0 || d.batch == 0) return GPB_OK;
    if (!d.C || (d.K > 0 && (!d.A || !d.B))) return GPB_ERR_INVALID;
    GemmParams p;
    p.M = d.M; p.N = d.N; p.K = d.K;
    p.A = d.A; p.lda = d.lda; p.B = d.B; p.ldb = d.ldb; p.C = d.C; p.ldc = d.ldc;
    p.alpha = d.alpha; p.beta = d.beta;
    p.mask = d.mask; p.mask_row0 = d.mask_row0; p.mask_col0 = d.mask_col0;
    p.mask_nb = d.mask_nb > 0 ? d.mask_nb : 1;
    p.krange = d.krange; p.kr_off = d.kr_off;
    p.strideA = d.strideA; p.strideB = d.strideB; p.strideC = d.strideC;
    p.batch = d.batch;
    p.tiles_m = p.tiles_n = 0;  // set per tile shape in dispatch_layouts
    auto al16 = [](const void* ptr, int64_t ld, int64_t stride) {
        return ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld % 2 == 0) && (stride % 2 == 0);
    };
    p.a_vec16 = al16(d.A, d.lda, d.strideA);
    p.b_vec16 = al16(d.B, d.ldb, d.strideB);
    p.c_vec16 = al16(d.C, d.ldc, d.strideC);
    cudaStream_t st = to_stream(s);
    // measured on B200 (scripts/gemm_bench.py): 4 warps x (32x64) 34.7 TF/s, 4 x (64x32) 34.7, 8 x (32x32) 33.2
    // Tile configurations measured on B200 (scripts/gemm_bench.py, 8192^3 / SYRK-lower 44.5k^2 K=512, TF/s):
    //   4 warps x (32x64), padded smem, 3 stages, 2 CTAs/SM : 34.7 / 33.2   <- default
    //   4 warps x (32x64), XOR-swizzled smem, 4 stages, 2/SM : 34.8 / 33.3   (variant 8, kept as a tuning hook)
    //   8 warps x (32x32), padded, 3 stages, 2 CTAs/SM       : 33.2 / 31.9   (variant 0, first version)
    //   rejected and removed: 4 x (64x32) (= default), swizzled 3 CTAs/SM at 168 regs (31.1, spills),
    //   128x128 CTA with 1 CTA/SM and 3 or 4 stages (31.9).
    //   variant 9 = operand staging through the TMA engine (cp.async.bulk row copies + mbarrier tx-count, SASS
    //   UBLKCP / SYNCS): correct, but 192 x 128-byte copies per stage are too fine-grained for it -- 20.1 TF/s for
    //   K-contiguous operands, 29.6 for the 512/1024-byte rows of the MN/MN layout -- so LDGSTS stays the default.
    //   (A 2-D tensor-map TMA cannot produce the padded rows, and its hardware swizzles are not conflict-free for
    //   the 8x4 f64 fragment: rows r and r^1 land in the same 32-byte bank group.)
    if (g_variant == 0) return dispatch_layouts<128, 64, 32, 32, 3, 2, false, false>(st, p, d);
    if (g_variant == 8) return dispatch_layouts<128, 64, 32, 64, 4, 2, true, false>(st, p, d);
    if (g_variant == 9) return dispatch_layouts<128, 64, 32, 64, 3, 2, false, true>(st, p, d);  // TMA bulk-copy staging
    // Few-tile products (the recursion of the diagonal blocks, the replicated M x M finishes, everything at N of a few thousand):
    // a 128 x 64 tile keeps one SM's DMMA pipe busy for ~1.1 us per 16-wide K step, so a launch of <= 16 such CTAs runs at the
    // speed of ONE SM's pipe for K / 16 steps (31-44 us each at K = 128..512, 56 % of a value + gradient at N = 1000:
    // profiles/r02_mll1000_launches.md) while the other SMs idle.  Quarter tiles (64 x 32, two warps of 32 x 32) put the same
    // product on four times as many SMs; used while the quarter-tile grid still fits one wave.  GPB_GEMM_SMALL_TILES=0 disables.
    if (g_small_tiles) {
        const int64_t t128 = ((d.M + 127) / 128) * ((d.N + 63) / 64) * (int64_t)d.batch;
        if (4 * t128 <= device_sm_count()) return dispatch_layouts<64, 32, 32, 32, 3, 4, false, false>(st, p, d);
    }
    return dispatch_layouts<128, 64, 32, 64, 3, 2, false, false>(st, p, d);
}

}  // namespace gpb
