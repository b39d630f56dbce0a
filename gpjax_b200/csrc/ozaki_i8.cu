// FP64 rank-k updates on the INT8 tensor pipe (Ozaki scheme) -- tcgen05.mma kind::i8, TMA-fed, TMEM accumulators.
//
// Why: every O(N^3) flop of the exact-GP path (reference call site: jnp.linalg.cholesky, gpjax/linalg/operations.py:54-55,
// and the reverse-mode solve it implies) is a rank-NB update C += alpha * A B^T.  The FP64 tensor instruction of sm_100a
// (DMMA.8x8x4) peaks at 37 TFLOP/s; the int8 tcgen05 pipe of the same chip is ~100x wider.  Splitting every fp64 operand
// row into `s` balanced radix-256 digits (the FULL int8 range, 8 bits per plane) after an exact power-of-two row scaling,
//      x_ik = 2^e_i * sum_p q^(p)_ik * 2^(-8 (p+1)),      -128 <= q <= 127,
// makes every digit-pair product an EXACT integer GEMM (int8 x int8 -> int32), and the fp64 result is recovered as
//      (A B^T)_ij ~= 2^(ea_i + eb_j) * sum_{t < s} 2^(-8 (t+2)) * sum_{p+q = t} (Q_a^(p) Q_b^(q)^T)_ij .
// All pairs of equal order t share one scale, so they are accumulated inside ONE int32 TMEM accumulator (|sum| <=
// (t+1) k 128^2 < 2^31 for s k < 2^17), i.e. s accumulator passes per output tile instead of s (s+1) / 2 separate GEMMs.
// The truncation error is bounded by the dropped orders: <= (s+1) 2^(-8 s - 2) 2^(ea_i + eb_j) k  (s = 6: 2^-47 per entry
// relative to the row maxima, s = 7: 2^-55 -- measured against the DMMA product in tests/test_gpu_ozaki.py and DESIGN
// section 12).  Round 1 used 7-bit digits (|q| <= 64): the same accuracy took one plane more, i.e. 28 instead of 22 (36
// instead of 28) digit-pair products.  An EVEN plane count also keeps the equal-plane pair (s/2, s/2) of order s: oz_has_diag.
//
// Kernel (persistent, one CTA per SM, 12 warps, warp-specialised; role loops are warp-uniform, elect.sync picks the issuing lane):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of 128 x 128-byte digit tiles (SWIZZLE_128B) into a ring of (A, B) slots
//   warp 1      MMA issuer  : tcgen05.mma.cta_group::1.kind::i8 (M=128, N=128, K=32 per instruction),
//                             tcgen05.commit releases ring slots / publishes accumulators through mbarriers
//   warp 2      TMEM allocator (512 columns = 4 accumulator stages of 128 x 128 int32)
//   warps 4-11  epilogue    : tcgen05.ld the int32 accumulator of order t, convert exactly to fp64, scale by 2^(-8 (t+2)) and
//                             add into per-thread fp64 registers (64 per thread) while the MMA warp already works on
//                             the next orders; after the last order: C (+)= alpha 2^(ea_i + eb_j) * acc, masked, ONE
//                             read-modify-write of the fp64 tile in HBM.
// Schedule (default): orders are processed in PAIRS (t, t+1) with two live accumulators -- ring slot i of a K-block holds
// (A_i, B_{t+1-i}); when it lands, acc_{t+1} += A_i B_{t+1-i}^T and acc_t += A_{i-1} B_{t+1-i}^T (A_{i-1} is still in the
// previous slot), so t+2 slot loads feed 2t+3 digit-pair products (16 instead of 28 loads per K-block at 7 planes).  That
// matters because the unpaired schedule saturates the L2 -> SM path (12.9 TB/s measured).  GPB_OZ_PAIR=0 selects the
// unpaired schedule (3 slots of 2 K-blocks), GPB_OZ_KERNEL=2 the plane-resident variant further below, GPB_OZ_NOLOAD=1
// disables the TMA copies (MMA pacing measurement) -- all three are measurement hooks, see profiles/r01_ozaki.md.
// The digit tiles are addressed directly in the [rows, s*k] digit matrix by TMA coordinates, so A and B may be the same
// buffer (SYRK) and no order-reversed copy exists.  Companion kernels: ozaki_slice_kernel (row digits), ozaki_slice_t_kernel +
// col_absmax_kernel (column digits for products that contract over the rows), col_wsum_* (augmented SGPR rows).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace gpb {
namespace {

constexpr int OZ_BM = 128, OZ_BN = 128, OZ_BK = 128;  // BK in bytes == int8 elements == one 128-byte swizzle row
constexpr int OZ_KBS_MAX = 2;                            // 128-byte K-blocks per ring stage (2 when the plane has an even count)
constexpr int OZ_STAGES = 3;
constexpr int OZ_MAX_RING = 6;                           // ring slots: 3 x 64 KB (2 K-blocks each) or 6 x 32 KB (paired-order schedule)
constexpr int OZ_KB_BYTES = (OZ_BM + OZ_BN) * OZ_BK;     // 32 KB: one (A, B) K-block pair
constexpr int OZ_STAGE_BYTES = OZ_KBS_MAX * OZ_KB_BYTES;  // 64 KB
constexpr int OZ_ACC_STAGES = 4;                         // x 128 TMEM columns
constexpr int OZ_THREADS = 384;
constexpr int OZ_EPI_WARP0 = 4, OZ_EPI_WARPS = 8;
constexpr int OZ_CHUNK_TILES = 32;  // output tile columns walked together so their B digits stay L2-resident
constexpr int OZ_SMEM_BYTES = OZ_STAGES * OZ_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
// OZ_BETA (= OZ_DIGIT_BITS = 8), oz_row_exponent and oz_fixed_point live in common.cuh (shared with the fused Gram -> digits kernel)

struct OzParams {
    int m, n;             // output extents
    int kblocks;          // k / 128 per digit plane
    int nslices;          // digit planes used (1 in raw mode)
    int mask;             // 0 none; 1 lower: (row0+i) >= (col0+j); 2 block-strict-upper: (row0+i)/mask_nb < (col0+j)/mask_nb;
                          // 3 block-upper: blocks with row block <= column block are live, and the DIAGONAL blocks go to C2 (v3 only)
    long long row0, col0, mask_nb;
    double* C2;           // mask 3: diagonal block b is the dense mask_nb x mask_nb matrix at C2 + b * mask_nb^2 (row stride mask_nb)
    int krange;           // v3 only: structural zeros of a triangular operand are never multiplied (same meaning as GemmDesc::krange)
    long long kr_off;
    double* C;            // fp64 in/out (MODE 1)
    long long ldc;
    int* Ci;              // int32 out (MODE 0)
    long long ldci;
    const double* sa;     // 2^ea_i
    const double* sb;     // 2^eb_j
    double alpha;
    int ntm, ntn;
    int kbs;              // K-blocks per ring stage (v1): 1 or 2
    int beta0;            // C = alpha A B^T instead of C += (C is never read)
    int pstride;          // digits between consecutive planes of one row (>= kblocks * 128)
    int nstages, stage_bytes;  // ring geometry (v1)
    int pair;             // v1: accumulate orders (t, t+1) together so every loaded A tile feeds two MMAs (see kernel)
    int noload;           // measurement hook (GPB_OZ_NOLOAD=1): the producer signals `full` without issuing TMA -> pure MMA pacing
    int bn;               // output tile width of the launched kernel variant (128: v1 / v3, 64: v2)
    int bm;               // output tile height the walk enumerates (128: v1 / v2; 256: v3, one tile per CTA pair)
    const int* planes_dev;  // optional device word overriding nslices (1..nslices): the conditioning guard, read in-kernel
    int max_ctas;         // > 0: launch at most this many CTAs (SMs left free for a concurrent look-ahead chain)
    int fx_bits;          // v4: fractional bits F of the int64 fixed-point recombination (see oz_fx_bits)
    int use_red;          // v4: C += v through red.global.add.f64 (no load in the epilogue); 0: load / add / store
};
// v4 recombines the orders EXACTLY in a 64-bit integer instead of an fp64 FMA chain: acc = sum_t P_t 2^(F - 8 t), one unit =
// 2^(-F - 16) 2^(ea + eb).  |P_t| <= (t + 1) K 2^14, so order 0 dominates and F = 61 - ceil(log2(K 2^14)) keeps |acc| < 2^62; orders
// with 8 t > F are rounded to the unit (2^-54 of the scale product at K = 1024: below the fp64 rounding of the result).  Why: one
// rounding per output entry instead of one per order, and one IMAD.WIDE per value on the integer pipe instead of a DADD + DFMA
// pair on the FP64 pipe the write-out also needs (profiles/r02_ozaki_epilogue.md; the time of the kernel did not change with it --
// what paced the epilogue was the load of C and of the column scales in the write-out, see oz_store_row_fx).
__host__ __device__ __forceinline__ int oz_fx_bits(long long K) {
    int lg = 0;
    while ((1ll << lg) < K * OZ_DIGIT_SQ_MAX) ++lg;
    int F = 61 - lg;
    if ((F & 7) == 7) --F;  // no order may need the multiplier 2^31 (not an int32)
    return F;
}
// planes actually used by this launch (uniform across the grid: every role of every CTA reads the same word)
__device__ __forceinline__ int oz_groups(const OzParams& p, int mode) {
    if (mode == 0) return 1;
    if (!p.planes_dev) return p.nslices;
    const int v = __ldg(p.planes_dev);
    return (v < 1 || v > p.nslices) ? p.nslices : v;  // a word nobody set (zero / garbage) means every plane, never fewer
}

// Even plane count s: the truncation set {p + q < s} would drop the pair (s/2, s/2), the product of two digits of EQUAL weight.
// In the rank-k updates of a factorisation the k-th entries of two panel rows have similar magnitude (column k of the panel has a
// characteristic scale), so their LEADING digits sit in the same plane and that pair is not noise: for a positive-definite update it
// is a coherent, same-signed term of relative size 2^(-8 s) per entry that adds up over K and over the block steps (measured:
// 14x the residual of the factor, 40x the MLL gradient error at cond 3e6: profiles/r02_radix256.md).  One extra product of
// order s restores the error level of an odd count; all other dropped pairs multiply a leading digit by a trailing (sign-random) one.
__device__ __host__ __forceinline__ bool oz_has_diag(int groups) { return groups >= 2 && (groups & 1) == 0; }

// K-blocks [kb0, kb1) a tile has to visit: the structural zeros of a triangular operand are skipped (v3).  All roles call this
// with the same arguments, so producer, issuer and epilogue stay in lock step.  n0 / m0: first column / row of the tile.
__device__ __forceinline__ void oz_krange(const OzParams& p, long long m0, long long n0, int bm, int& kb0, int& kb1) {
    kb0 = 0;
    kb1 = p.kblocks;
    if (p.krange == KR_B_LOWER) {         // B(n,k) == 0 for k > n + off
        long long e = (n0 + OZ_BN - 1 + p.kr_off) / OZ_BK + 1;
        kb1 = e < 1 ? 1 : (e < p.kblocks ? (int)e : p.kblocks);
    } else if (p.krange == KR_B_UPPER) {  // B(n,k) == 0 for k < n + off
        long long b = (n0 + p.kr_off) / OZ_BK;
        kb0 = b < 0 ? 0 : (b < p.kblocks ? (int)b : p.kblocks - 1);
    } else if (p.krange == KR_A_LOWER) {  // A(m,k) == 0 for k > m + off
        long long e = (m0 + bm - 1 + p.kr_off) / OZ_BK + 1;
        kb1 = e < 1 ? 1 : (e < p.kblocks ? (int)e : p.kblocks);
    } else if (p.krange == KR_A_UPPER) {  // A(m,k) == 0 for k < m + off
        long long b = (m0 + p.kr_off) / OZ_BK;
        kb0 = b < 0 ? 0 : (b < p.kblocks ? (int)b : p.kblocks - 1);
    }
}

// ---- tile enumeration shared by the three roles: column chunks -> tile rows -> tile columns, dead tiles of the lower
// mask never enumerated ----------------------------------------------------------------------------------------------
struct TileWalk {
    int chunk = 0, tm = 0, c0 = 0, c1 = 0;
    long long base = 0;  // linear index of the first tile of (chunk, tm)
    __device__ int live_end(const OzParams& p, int tm_) const {  // one past the last live tile column of tile row tm_
        if (p.mask != 1) return p.ntn;
        long long last_row = p.row0 + (long long)tm_ * p.bm + p.bm - 1;
        long long d = last_row - p.col0;
        if (d < 0) return 0;
        long long e = d / p.bn + 1;
        return e < p.ntn ? (int)e : p.ntn;
    }
    __device__ int live_begin(const OzParams& p, int tm_) const {  // first live tile column of tile row tm_
        if (p.mask != 2 && p.mask != 3) return 0;
        long long rb = (p.row0 + (long long)tm_ * p.bm) / p.mask_nb;     // block of the tile's FIRST row (smallest)
        // mask 2: col0 + tn*BN + BN-1 >= (rb+1)*nb;  mask 3 (diagonal blocks live too): ... >= rb*nb
        long long need = (rb + (p.mask == 2 ? 1 : 0)) * p.mask_nb - (p.bn - 1) - p.col0;
        if (need <= 0) return 0;
        long long b = (need + p.bn - 1) / p.bn;
        return b < p.ntn ? (int)b : p.ntn;
    }
    __device__ int lo(const OzParams& p) const {
        int b = live_begin(p, tm);
        return b > c0 ? b : c0;
    }
    __device__ int count(const OzParams& p) const {
        int e = live_end(p, tm);
        int hi = e < c1 ? e : c1;
        int l = lo(p);
        return hi > l ? hi - l : 0;
    }
    __device__ void set_chunk(const OzParams& p, int ch) {
        chunk = ch;
        c0 = ch * OZ_CHUNK_TILES;
        c1 = c0 + OZ_CHUNK_TILES < p.ntn ? c0 + OZ_CHUNK_TILES : p.ntn;
        tm = 0;
    }
    __device__ void init(const OzParams& p) { set_chunk(p, 0); base = 0; }
    // advance to the tile with linear index idx (monotonically increasing calls); false when past the end
    __device__ bool seek(const OzParams& p, long long idx, int& tm_out, int& tn_out) {
        const int nchunks = (p.ntn + OZ_CHUNK_TILES - 1) / OZ_CHUNK_TILES;
        while (true) {
            if (chunk >= nchunks) return false;
            int cnt = count(p);
            if (idx < base + cnt) {
                tm_out = tm;
                tn_out = lo(p) + (int)(idx - base);
                return true;
            }
            base += cnt;
            if (++tm >= p.ntm) set_chunk(p, chunk + 1);
        }
    }
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (-> launch error) instead of hanging the device
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
    if (mbar_try(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// one lane of a CONVERGED warp; keeping the role loops warp-uniform lets the compiler hold descriptors / addresses in uniform
// registers (an `if (lane == 0)` role makes every UTCIMMA / UTMALDG operand pass through an ELECT + R2UR waterfall loop)
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_i8(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void tc_ld16(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// K-major operand tile in shared memory: rows of 128 bytes, 8-row groups 1024 bytes apart, 128-byte swizzle
// (what TMA SWIZZLE_128B writes for a {128 B, rows} box into a 1024-byte aligned buffer).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(unsigned saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address, 16-byte units           bits [0,14)
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused: swizzled)  bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows x 128 B      bits [32,46)
    d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)             bits [46,48)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B                            bits [61,64)
    return d;
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 128, M = 128
constexpr unsigned OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(OZ_BN >> 3) << 17) | ((unsigned)(OZ_BM >> 4) << 24);

__device__ __forceinline__ double exact_i2d(int v) {  // exact int32 -> fp64 on the FP64 add pipe (no I2F)
    return __hiloint2double(0x43300000, (int)((unsigned)v ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
}

// MODE 0: Ci = A B^T (raw int32, test / building block); MODE 1: C += alpha * 2^(ea+eb) * sum_t 2^(-8(t+2)) P_t
template <int MODE>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzParams p) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OZ_STAGES * OZ_STAGE_BYTES);
    uint64_t* full = bars;                         // [<= 6]     TMA -> MMA
    uint64_t* empty = bars + OZ_MAX_RING;          // [<= 6]     MMA -> TMA
    uint64_t* tfull = bars + 2 * OZ_MAX_RING;      // [ACC]      MMA -> epilogue
    uint64_t* tempty = tfull + OZ_ACC_STAGES;      // [ACC]      epilogue -> MMA
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + OZ_ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < OZ_MAX_RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < OZ_ACC_STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], OZ_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    const int groups = oz_groups(p, MODE);

    if (warp == 0) {
        {  // ===== TMA producer (whole warp walks the schedule, one elected lane issues) =====
            TileWalk w; w.init(p);
            int stage = 0; unsigned phase = 0;
            int tm, tn;
            for (long long idx = blockIdx.x; w.seek(p, idx, tm, tn); idx += gridDim.x) {
                const int m0 = tm * OZ_BM, n0 = tn * OZ_BN;
                for (int t = 0; t < groups;) {
                    if (p.pair && ((groups - t) & 1) == 0) {  // odd plane count: the cheapest order (t = 0) is the unpaired one
                        // orders (t, t+1) together: stage i of a K-block holds (A_i, B_{t+1-i}), i = 0..t+1
                        for (int kb = 0; kb < p.kblocks; ++kb) {
                            for (int i = 0; i <= t + 1; ++i) {
                                mbar_wait_bounded(&empty[stage], phase ^ 1u);
                                uint8_t* sA = smem + stage * p.stage_bytes;
                                if (elect_one()) {
                                    if (p.noload) {
                                        mbar_arrive(&full[stage]);
                                    } else {
                                        mbar_arrive_expect_tx(&full[stage], OZ_KB_BYTES);
                                        tma_load_2d(sA, &tmA, &full[stage], i * p.pstride + kb * OZ_BK, m0);
                                        tma_load_2d(sA + OZ_BM * OZ_BK, &tmB, &full[stage], (t + 1 - i) * p.pstride + kb * OZ_BK, n0);
                                    }
                                }
                                __syncwarp();
                                if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
                            }
                        }
                        t += 2;
                        continue;
                    }
                    for (int pa = 0; pa <= t; ++pa) {
                        const int xa0 = pa * p.pstride, xb0 = (t - pa) * p.pstride;
                        for (int kb = 0; kb < p.kblocks; kb += p.kbs) {
                            mbar_wait_bounded(&empty[stage], phase ^ 1u);
                            uint8_t* sA = smem + stage * p.stage_bytes;
                            if (elect_one()) {
                                if (p.noload) {
                                    mbar_arrive(&full[stage]);
                                } else {
                                    mbar_arrive_expect_tx(&full[stage], p.kbs * OZ_KB_BYTES);
                                    for (int j = 0; j < p.kbs; ++j) {
                                        tma_load_2d(sA + j * OZ_KB_BYTES, &tmA, &full[stage], xa0 + (kb + j) * OZ_BK, m0);
                                        tma_load_2d(sA + j * OZ_KB_BYTES + OZ_BM * OZ_BK, &tmB, &full[stage], xb0 + (kb + j) * OZ_BK, n0);
                                    }
                                }
                            }
                            __syncwarp();
                            if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
                        }
                    }
                    ++t;
                }
            }
        }
    } else if (warp == 1) {
        {  // ===== MMA issuer (whole warp walks the schedule, one elected lane issues) =====
            TileWalk w; w.init(p);
            int stage = 0; unsigned phase = 0;
            int acc = 0; unsigned aphase = 0;
            int tm, tn;
            for (long long idx = blockIdx.x; w.seek(p, idx, tm, tn); idx += gridDim.x) {
                for (int t = 0; t < groups;) {
                    if (p.pair && ((groups - t) & 1) == 0) {  // odd plane count: the cheapest order (t = 0) is the unpaired one
                        // orders (t, t+1) in two accumulators: when stage i = (A_i, B_{t+1-i}) lands,
                        //   acc_hi += A_i B_{t+1-i}^T   and   acc_lo += A_{i-1} B_{t+1-i}^T  (A_{i-1} still sits in the previous stage),
                        // so t+2 stage loads feed 2t+3 digit-pair products instead of 2t+3 loads.
                        const int a_lo = acc; const unsigned ph_lo = aphase;
                        if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
                        const int a_hi = acc; const unsigned ph_hi = aphase;
                        if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
                        mbar_wait_bounded(&tempty[a_lo], ph_lo ^ 1u);
                        mbar_wait_bounded(&tempty[a_hi], ph_hi ^ 1u);
                        tc_fence_after();
                        const unsigned d_lo = tmem_base + (unsigned)(a_lo * OZ_BN), d_hi = tmem_base + (unsigned)(a_hi * OZ_BN);
                        int prev = 0;
                        for (int kb = 0; kb < p.kblocks; ++kb) {
                            for (int i = 0; i <= t + 1; ++i) {
                                mbar_wait_bounded(&full[stage], phase);
                                tc_fence_after();
                                const unsigned sA = smem_u32(smem + stage * p.stage_bytes);
                                const unsigned sP = smem_u32(smem + prev * p.stage_bytes);
                                if (elect_one()) {
                                    const uint64_t da = umma_desc_k_sw128(sA), db = umma_desc_k_sw128(sA + OZ_BM * OZ_BK);
#pragma unroll
                                    for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                        tc_mma_i8(d_hi, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), OZ_IDESC, (kb | i | kk) != 0);
                                    if (i >= 1) {
                                        const uint64_t dp = umma_desc_k_sw128(sP);
#pragma unroll
                                        for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                            tc_mma_i8(d_lo, dp + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), OZ_IDESC,
                                                      !(kb == 0 && i == 1 && kk == 0));
                                        tc_commit(&empty[prev]);  // A_{i-1} and its B are done
                                    }
                                    if (i == t + 1) {
                                        tc_commit(&empty[stage]);  // last stage of this K-block: nothing pairs with A_{t+1} later
                                        if (kb == p.kblocks - 1) { tc_commit(&tfull[a_lo]); tc_commit(&tfull[a_hi]); }
                                    }
                                }
                                __syncwarp();
                                prev = stage;
                                if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
                            }
                        }
                        t += 2;
                        continue;
                    }
                    mbar_wait_bounded(&tempty[acc], aphase ^ 1u);
                    tc_fence_after();
                    const unsigned d_tmem = tmem_base + (unsigned)(acc * OZ_BN);
                    const int nkb = (t + 1) * p.kblocks;
                    for (int kb = 0; kb < nkb; kb += p.kbs) {
                        mbar_wait_bounded(&full[stage], phase);
                        tc_fence_after();
                        const unsigned sA = smem_u32(smem + stage * p.stage_bytes);
                        if (elect_one()) {
                            for (int j = 0; j < p.kbs; ++j) {
                                const uint64_t da = umma_desc_k_sw128(sA + j * OZ_KB_BYTES);
                                const uint64_t db = umma_desc_k_sw128(sA + j * OZ_KB_BYTES + OZ_BM * OZ_BK);
#pragma unroll
                                for (int kk = 0; kk < OZ_BK / 32; ++kk)  // +32 bytes along K inside the swizzle row = +2 in the address field
                                    tc_mma_i8(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), OZ_IDESC, (kb | j | kk) != 0);
                            }
                            tc_commit(&empty[stage]);  // frees the ring slot once these MMAs have read it
                            if (kb + p.kbs >= nkb) tc_commit(&tfull[acc]);  // accumulator of order t complete
                        }
                        __syncwarp();
                        if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
                    }
                    if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
                    ++t;
                }
            }
        }
    } else if (warp >= OZ_EPI_WARP0) {
        // ===== epilogue: warp (4 + 4 h + q) owns TMEM lanes [32 q, 32 q + 32) and tile columns [64 h, 64 h + 64) =====
        const int q = warp & 3, h = (warp - OZ_EPI_WARP0) >> 2;
        TileWalk w; w.init(p);
        int acc = 0; unsigned aphase = 0;
        int tm, tn;
        for (long long idx = blockIdx.x; w.seek(p, idx, tm, tn); idx += gridDim.x) {
            const int row = tm * OZ_BM + q * 32 + lane;
            const int col0 = tn * OZ_BN + h * 64;
            double accd[MODE == 1 ? 64 : 1];
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 64; ++j) accd[j] = 0.0;
            }
            for (int t = 0; t < groups; ++t) {
                mbar_wait_bounded(&tfull[acc], aphase);
                tc_fence_after();
                const double sc = __hiloint2double((1023 - OZ_BETA * (t + 2)) << 20, 0);  // 2^(-8 (t+2))
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    unsigned r[32];
                    tc_ld32(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(acc * OZ_BN + h * 64 + c * 32), r);
                    if (MODE == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) accd[c * 32 + j] = fma(exact_i2d((int)r[j]), sc, accd[c * 32 + j]);
                    } else {
                        if (row < p.m) {
                            int* dst = p.Ci + (long long)row * p.ldci + col0 + c * 32;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + c * 32 + j < p.n) dst[j] = (int)r[j];
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
            }
            if (MODE == 1 && row < p.m) {
                const double sr = p.alpha * __ldg(p.sa + row);
                double* crow = p.C + (long long)row * p.ldc;
                const long long grow = p.row0 + row;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int col = col0 + j;
                    bool live = col < p.n;
                    if (p.mask == 1) live = live && (grow >= p.col0 + col);
                    else if (p.mask == 2) live = live && (grow / p.mask_nb < (p.col0 + col) / p.mask_nb);
                    if (live) {
                        const double v = accd[j] * (sr * __ldg(p.sb + col));
                        crow[col] = p.beta0 ? v : crow[col] + v;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- variant 2: plane-resident schedule ------------------------------------------------------------------------------
// v1 above re-reads every digit tile once per order: 28 (A,B) tile pairs x 32 KB per K-block of a 128 x 128 output tile,
// which saturates the L2 -> SM path (measured 12.9 TB/s == the chip's LTS cap) at ~1.7 Pop/s.  Here ALL `s` order
// accumulators of a 128 x 64 output tile are live in TMEM at once (s x 64 columns <= 512), the loop over the K-blocks is
// outermost, and per K-block each digit plane of A (16 KB) and of B (8 KB) is loaded exactly once and reused by every
// pair (p, q), p + q < s: 168 KB instead of 896 KB per K-block for 7 planes (per output element: 2.7x less L2 traffic).
//   smem: A ring of 6 x 16 KB (plane tiles stream through it in consumption order) + B planes double-buffered by K-block
//         parity, 2 x 8 x 8 KB; every buffer has its own full/empty mbarrier pair, tcgen05.commit frees a buffer as soon
//         as the last MMA reading it has retired (A_p after its q sweep, B_q after the sweep of p = s-1-q).
//   TMEM: accumulator of order t at columns [64 t, 64 t + 64); the epilogue folds the orders into 32 fp64 registers per
//         thread, releases TMEM, then does the single read-modify-write of C while the next tile's MMAs already run.
constexpr int O2_BN = 64;
constexpr int O2_ASTAGES = 6;
constexpr int O2_BBUFS = 16;
constexpr int O2_A_BYTES = OZ_BM * OZ_BK;  // 16 KB
constexpr int O2_B_BYTES = O2_BN * OZ_BK;  //  8 KB
constexpr int O2_DATA_BYTES = O2_ASTAGES * O2_A_BYTES + O2_BBUFS * O2_B_BYTES;  // 224 KB
constexpr int O2_SMEM_BYTES = O2_DATA_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
constexpr unsigned O2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(O2_BN >> 3) << 17) | ((unsigned)(OZ_BM >> 4) << 24);
static_assert(O2_SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

template <int MODE>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_i8_kernel_v2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzParams p) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint8_t* sBbase = smem + O2_ASTAGES * O2_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + O2_DATA_BYTES);
    uint64_t* afull = bars;                              // [6]
    uint64_t* aempty = bars + O2_ASTAGES;                // [6]
    uint64_t* bfull = bars + 2 * O2_ASTAGES;             // [16]
    uint64_t* bempty = bfull + O2_BBUFS;                 // [16]
    uint64_t* tfull = bempty + O2_BBUFS;                 // [1]
    uint64_t* tempty = tfull + 1;                        // [1]
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < O2_ASTAGES; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < O2_BBUFS; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        mbar_init(tfull, 1);
        mbar_init(tempty, OZ_EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    const int S = oz_groups(p, MODE);
    const int KB = p.kblocks;

    if (warp == 0) {
        {  // ===== TMA producer: per K-block A_0, B_0..B_{S-1}, A_1, ..., A_{S-1} (consumption order) =====
            TileWalk w; w.init(p);
            int astage = 0; unsigned aphase = 0; unsigned kbc = 0;
            int tm, tn;
            for (long long idx = blockIdx.x; w.seek(p, idx, tm, tn); idx += gridDim.x) {
                const int m0 = tm * OZ_BM, n0 = tn * O2_BN;
                for (int kb = 0; kb < KB; ++kb, ++kbc) {
                    const int par = (int)(kbc & 1u);
                    const unsigned bph = (kbc >> 1) & 1u;
                    for (int pl = 0; pl < S; ++pl) {
                        mbar_wait_bounded(&aempty[astage], aphase ^ 1u);
                        if (elect_one()) {
                            if (p.noload) {
                                mbar_arrive(&afull[astage]);
                            } else {
                                mbar_arrive_expect_tx(&afull[astage], O2_A_BYTES);
                                tma_load_2d(smem + astage * O2_A_BYTES, &tmA, &afull[astage], pl * p.pstride + kb * OZ_BK, m0);
                            }
                        }
                        __syncwarp();
                        if (++astage == O2_ASTAGES) { astage = 0; aphase ^= 1u; }
                        if (pl == 0) {
                            for (int q = 0; q < S; ++q) {
                                const int buf = par * 8 + q;
                                mbar_wait_bounded(&bempty[buf], bph ^ 1u);
                                if (elect_one()) {
                                    if (p.noload) {
                                        mbar_arrive(&bfull[buf]);
                                    } else {
                                        mbar_arrive_expect_tx(&bfull[buf], O2_B_BYTES);
                                        tma_load_2d(sBbase + buf * O2_B_BYTES, &tmB, &bfull[buf], q * p.pstride + kb * OZ_BK, n0);
                                    }
                                }
                                __syncwarp();
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        {  // ===== MMA issuer =====
            TileWalk w; w.init(p);
            int astage = 0; unsigned aphase = 0; unsigned kbc = 0; unsigned tph = 0;
            int tm, tn;
            for (long long idx = blockIdx.x; w.seek(p, idx, tm, tn); idx += gridDim.x) {
                mbar_wait_bounded(tempty, tph ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < KB; ++kb, ++kbc) {
                    const int par = (int)(kbc & 1u);
                    const unsigned bph = (kbc >> 1) & 1u;
                    for (int pl = 0; pl < S; ++pl) {
                        mbar_wait_bounded(&afull[astage], aphase);
                        const uint64_t da = umma_desc_k_sw128(smem_u32(smem + astage * O2_A_BYTES));
                        for (int q = 0; q < S - pl; ++q) {
                            const int buf = par * 8 + q;
                            if (pl == 0) mbar_wait_bounded(&bfull[buf], bph);
                            tc_fence_after();
                            const uint64_t db = umma_desc_k_sw128(smem_u32(sBbase + buf * O2_B_BYTES));
                            const unsigned d_tmem = tmem_base + (unsigned)((pl + q) * O2_BN);
                            if (elect_one()) {
#pragma unroll
                                for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                    tc_mma_i8(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), O2_IDESC,
                                              !(kb == 0 && pl == 0 && kk == 0));
                            }
                            __syncwarp();
                        }
                        if (elect_one()) {
                            tc_commit(&aempty[astage]);                  // A_pl fully consumed
                            tc_commit(&bempty[par * 8 + (S - 1 - pl)]);  // B_{S-1-pl} had its last reader in this sweep
                            if (kb == KB - 1 && pl == S - 1) tc_commit(tfull);
                        }
                        __syncwarp();
                        if (++astage == O2_ASTAGES) { astage = 0; aphase ^= 1u; }
                    }
                }
                tph ^= 1u;
            }
        }
    } else if (warp >= OZ_EPI_WARP0) {
        // ===== epilogue: warp (4 + 4 h + q4) owns TMEM lanes [32 q4, 32 q4 + 32) and tile columns [32 h, 32 h + 32) =====
        const int q4 = warp & 3, h = (warp - OZ_EPI_WARP0) >> 2;
        TileWalk w; w.init(p);
        unsigned eph = 0;
        int tm, tn;
        for (long long idx = blockIdx.x; w.seek(p, idx, tm, tn); idx += gridDim.x) {
            const int row = tm * OZ_BM + q4 * 32 + lane;
            const int col0 = tn * O2_BN + h * 32;
            double accd[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) accd[j] = 0.0;
            mbar_wait_bounded(tfull, eph);
            tc_fence_after();
            for (int t = 0; t < S; ++t) {
                unsigned r[32];
                tc_ld32(tmem_base + ((unsigned)(q4 * 32) << 16) + (unsigned)(t * O2_BN + h * 32), r);
                if (MODE == 1) {
                    const double sc = __hiloint2double((1023 - OZ_BETA * (t + 2)) << 20, 0);  // 2^(-8 (t+2))
#pragma unroll
                    for (int j = 0; j < 32; ++j) accd[j] = fma(exact_i2d((int)r[j]), sc, accd[j]);
                } else if (row < p.m) {
                    int* dst = p.Ci + (long long)row * p.ldci + col0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.n) dst[j] = (int)r[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
            eph ^= 1u;
            if (MODE == 1 && row < p.m) {
                const double sr = p.alpha * __ldg(p.sa + row);
                double* crow = p.C + (long long)row * p.ldc;
                const long long grow = p.row0 + row;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = col0 + j;
                    bool live = col < p.n;
                    if (p.mask == 1) live = live && (grow >= p.col0 + col);
                    else if (p.mask == 2) live = live && (grow / p.mask_nb < (p.col0 + col) / p.mask_nb);
                    if (live) {
                        const double v = accd[j] * (sr * __ldg(p.sb + col));
                        crow[col] = p.beta0 ? v : crow[col] + v;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- variant 3 (default): CTA pairs, tcgen05.mma.cta_group::2, M = 256 x N = 128 per pair -----------------------------------
// Measured on v1 (profiles/r01_ozaki.md section 3, profiles/r02_ozaki_cg2.md): with the TMA loads switched off an M128 x N128 x K32
// kind::i8 instruction still paces at ~145 clk, an N = 64 one at ~110 clk, cuBLASLt's N = 256 ones at ~210 clk, i.e.
// ~(4096 + 32 N) / 56 clk: the instruction is paced by the 32 (M + N) operand bytes it pulls out of shared memory (~56 B/clk),
// not by the int8 pipe (64 clk for 128 x 128 x 32).  A CTA pair sharing one MMA halves the B bytes per SM: each CTA stages its
// own 128 rows of A (16 KB per K-block) and HALF of the B tile (64 rows, 8 KB); one elected thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = 128), each SM accumulates its 128 x 128 half in its own TMEM.  Per SM that is 6 KB
// instead of 8 KB of operand reads per instruction and 24 KB instead of 32 KB of L2 -> SM traffic per ring slot.
//   * cluster of 2 CTAs (__cluster_dims__), one cluster per TPC, persistent over 256 x 128 output tiles (same TileWalk);
//   * TMA: cp.async.bulk.tensor.2d.cta_group::2, both CTAs complete_tx on the LEADER's `full` barrier (mapa to rank 0);
//   * tcgen05.commit.cta_group::2 ... multicast::cluster frees the ring slot / publishes the accumulator in BOTH CTAs;
//   * the epilogue warps of both CTAs arrive on the leader's `tempty` barrier (count 2 x 8 warps);
//   * same paired-order schedule as v1 (ring slot i of a K-block = (A_i, B_{t+1-i}), two live accumulators), 8-slot ring.
constexpr int C2_BM = 256;                        // rows per pair tile (128 per CTA)
constexpr int C2_BNH = OZ_BN / 2;                 // B rows staged by each CTA
constexpr int C2_A_BYTES = OZ_BM * OZ_BK;         // 16 KB
constexpr int C2_B_BYTES = C2_BNH * OZ_BK;        //  8 KB
constexpr int C2_SLOT_BYTES = C2_A_BYTES + C2_B_BYTES;  // 24 KB
constexpr int C2_RING = 8;
constexpr int C2_SMEM_BYTES = C2_RING * C2_SLOT_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 128, M = 256 (cta_group::2)
constexpr unsigned C2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(OZ_BN >> 3) << 17) | ((unsigned)(C2_BM >> 4) << 24);
static_assert(C2_SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa_u32(const void* p, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// same without the release fence: the accumulator hand-back orders TMEM reads (tcgen05.wait::ld + fence::before_thread_sync), not
// global memory -- a release here would wait for every outstanding red.global of the previous tile's write-out
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are signalled on the mbarrier at the
// shared::cluster address `bar_cluster` (the leader's `full` barrier)
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, unsigned bar_cluster, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(bar_cluster), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tc_commit_cg2(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                     smem_u32(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_i8_cg2(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// fp64 write-out of one thread's 64 running sums (row `row`, tile columns [col0, col0 + 64)): C (+)= alpha 2^(ea+eb) acc, masked
__device__ __forceinline__ void oz_store_row(const OzParams& p, const double (&accd)[64], int row, int col0) {
    const double sr = p.alpha * __ldg(p.sa + row);
    const long long grow = p.row0 + row;
    double* crow = p.C + (long long)row * p.ldc;  // indexed by the local column
    long long cshift = 0;
    bool diag_blk = false;
    if (p.mask == 3) {  // a 64-column slab lies inside one column block (mask_nb is a multiple of 128)
        const long long rb = grow / p.mask_nb, cb = (p.col0 + col0) / p.mask_nb;
        diag_blk = rb == cb;
        if (diag_blk) {  // dense diagonal block rb of C2, addressed by (row, column) inside the block
            crow = p.C2 + rb * p.mask_nb * p.mask_nb + (grow - rb * p.mask_nb) * p.mask_nb;
            cshift = p.col0 - cb * p.mask_nb;
        }
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        const int col = col0 + j;
        bool live = col < p.n;
        if (p.mask == 1) live = live && (grow >= p.col0 + col);
        else if (p.mask == 2) live = live && (grow / p.mask_nb < (p.col0 + col) / p.mask_nb);
        else if (p.mask == 3) live = live && (diag_blk || grow / p.mask_nb < (p.col0 + col) / p.mask_nb);
        if (live) {
            const double v = accd[j] * (sr * __ldg(p.sb + col));
            double* dst = crow + col + cshift;
            *dst = p.beta0 ? v : *dst + v;
        }
    }
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1)
ozaki_i8_kernel_cg2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzParams p) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);  // identical offset in both CTAs of the pair
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C2_RING * C2_SLOT_BYTES);
    uint64_t* full = bars;                       // [8]   TMA (both CTAs) -> MMA; only the leader's copies are used
    uint64_t* empty = bars + C2_RING;            // [8]   MMA -> TMA, multicast to both CTAs
    uint64_t* tfull = bars + 2 * C2_RING;        // [ACC] MMA -> epilogue, multicast to both CTAs
    uint64_t* tempty = tfull + OZ_ACC_STAGES;    // [ACC] epilogue (both CTAs) -> MMA; only the leader's copies are used
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + OZ_ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C2_RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < OZ_ACC_STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * OZ_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 2) {  // the same warp of BOTH CTAs executes the pair allocation; each gets the base address in its own smem
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits + TMEM base visible to the peer before anything signals across the pair
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    const int groups = oz_groups(p, MODE);

    if (warp == 0) {
        {  // ===== TMA producer (both CTAs): own 128 rows of A, own 64 rows of B; bytes are counted on the leader's barrier =====
            TileWalk w; w.init(p);
            int stage = 0; unsigned phase = 0;
            int tm, tn;
            for (long long idx = pair; w.seek(p, idx, tm, tn); idx += npairs) {
                const int m0 = tm * C2_BM + (int)rank * OZ_BM, n0 = tn * OZ_BN + (int)rank * C2_BNH;
                int kb0, kb1;
                oz_krange(p, (long long)tm * C2_BM, (long long)tn * OZ_BN, C2_BM, kb0, kb1);
                for (int t = 0; t < groups;) {
                    const bool paired = p.pair && ((groups - t) & 1) == 0;  // odd plane count: order 0 runs unpaired
                    // paired: slot i of a K-block holds (A_i, B_{t+1-i}), i = 0..t+1; unpaired: (A_pa, B_{t-pa}), pa = 0..t
                    const int nsl = paired ? t + 2 : t + 1;
                    const int bsum = paired ? t + 1 : t;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        for (int i = 0; i < nsl; ++i) {
                            mbar_wait_bounded(&empty[stage], phase ^ 1u);
                            uint8_t* sA = smem + stage * C2_SLOT_BYTES;
                            if (elect_one()) {
                                if (p.noload) {
                                    if (rank == 0) mbar_arrive(&full[stage]);
                                } else {
                                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C2_SLOT_BYTES);
                                    const unsigned fb = mapa_u32(&full[stage], 0);
                                    tma_load_2d_cg2(sA, &tmA, fb, i * p.pstride + kb * OZ_BK, m0);
                                    tma_load_2d_cg2(sA + C2_A_BYTES, &tmB, fb, (bsum - i) * p.pstride + kb * OZ_BK, n0);
                                }
                            }
                            __syncwarp();
                            if (++stage == C2_RING) { stage = 0; phase ^= 1u; }
                        }
                    }
                    t += paired ? 2 : 1;
                }
                if (oz_has_diag(groups)) {  // even plane count: the equal-plane product A_h B_h^T of order `groups`
                    const int hp = groups >> 1;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait_bounded(&empty[stage], phase ^ 1u);
                        uint8_t* sA = smem + stage * C2_SLOT_BYTES;
                        if (elect_one()) {
                            if (p.noload) {
                                if (rank == 0) mbar_arrive(&full[stage]);
                            } else {
                                if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C2_SLOT_BYTES);
                                const unsigned fb = mapa_u32(&full[stage], 0);
                                tma_load_2d_cg2(sA, &tmA, fb, hp * p.pstride + kb * OZ_BK, m0);
                                tma_load_2d_cg2(sA + C2_A_BYTES, &tmB, fb, hp * p.pstride + kb * OZ_BK, n0);
                            }
                        }
                        __syncwarp();
                        if (++stage == C2_RING) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {  // ===== MMA issuer: leader CTA only, one elected lane =====
            TileWalk w; w.init(p);
            int stage = 0; unsigned phase = 0;
            int acc = 0; unsigned aphase = 0;
            int tm, tn;
            for (long long idx = pair; w.seek(p, idx, tm, tn); idx += npairs) {
                int kb0, kb1;
                oz_krange(p, (long long)tm * C2_BM, (long long)tn * OZ_BN, C2_BM, kb0, kb1);
                for (int t = 0; t < groups;) {
                    if (p.pair && ((groups - t) & 1) == 0) {
                        const int a_lo = acc; const unsigned ph_lo = aphase;
                        if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
                        const int a_hi = acc; const unsigned ph_hi = aphase;
                        if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
                        mbar_wait_bounded(&tempty[a_lo], ph_lo ^ 1u);
                        mbar_wait_bounded(&tempty[a_hi], ph_hi ^ 1u);
                        tc_fence_after();
                        const unsigned d_lo = tmem_base + (unsigned)(a_lo * OZ_BN), d_hi = tmem_base + (unsigned)(a_hi * OZ_BN);
                        int prev = 0;
                        for (int kb = kb0; kb < kb1; ++kb) {
                            for (int i = 0; i <= t + 1; ++i) {
                                mbar_wait_bounded(&full[stage], phase);
                                tc_fence_after();
                                const unsigned sA = smem_u32(smem + stage * C2_SLOT_BYTES);
                                const unsigned sP = smem_u32(smem + prev * C2_SLOT_BYTES);
                                if (elect_one()) {
                                    const uint64_t da = umma_desc_k_sw128(sA), db = umma_desc_k_sw128(sA + C2_A_BYTES);
#pragma unroll
                                    for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                        tc_mma_i8_cg2(d_hi, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), C2_IDESC,
                                                      ((kb - kb0) | i | kk) != 0);
                                    if (i >= 1) {
                                        const uint64_t dp = umma_desc_k_sw128(sP);
#pragma unroll
                                        for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                            tc_mma_i8_cg2(d_lo, dp + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), C2_IDESC,
                                                          !(kb == kb0 && i == 1 && kk == 0));
                                        tc_commit_cg2(&empty[prev]);
                                    }
                                    if (i == t + 1) {
                                        tc_commit_cg2(&empty[stage]);
                                        if (kb == kb1 - 1) { tc_commit_cg2(&tfull[a_lo]); tc_commit_cg2(&tfull[a_hi]); }
                                    }
                                }
                                __syncwarp();
                                prev = stage;
                                if (++stage == C2_RING) { stage = 0; phase ^= 1u; }
                            }
                        }
                        t += 2;
                        continue;
                    }
                    mbar_wait_bounded(&tempty[acc], aphase ^ 1u);
                    tc_fence_after();
                    const unsigned d_tmem = tmem_base + (unsigned)(acc * OZ_BN);
                    const int nkb = (t + 1) * (kb1 - kb0);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait_bounded(&full[stage], phase);
                        tc_fence_after();
                        const unsigned sA = smem_u32(smem + stage * C2_SLOT_BYTES);
                        if (elect_one()) {
                            const uint64_t da = umma_desc_k_sw128(sA), db = umma_desc_k_sw128(sA + C2_A_BYTES);
#pragma unroll
                            for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                tc_mma_i8_cg2(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), C2_IDESC, (kb | kk) != 0);
                            tc_commit_cg2(&empty[stage]);
                            if (kb + 1 >= nkb) tc_commit_cg2(&tfull[acc]);
                        }
                        __syncwarp();
                        if (++stage == C2_RING) { stage = 0; phase ^= 1u; }
                    }
                    if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
                    ++t;
                }
                if (oz_has_diag(groups)) {
                    mbar_wait_bounded(&tempty[acc], aphase ^ 1u);
                    tc_fence_after();
                    const unsigned d_tmem = tmem_base + (unsigned)(acc * OZ_BN);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait_bounded(&full[stage], phase);
                        tc_fence_after();
                        const unsigned sA = smem_u32(smem + stage * C2_SLOT_BYTES);
                        if (elect_one()) {
                            const uint64_t da = umma_desc_k_sw128(sA), db = umma_desc_k_sw128(sA + C2_A_BYTES);
#pragma unroll
                            for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                tc_mma_i8_cg2(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), C2_IDESC, ((kb - kb0) | kk) != 0);
                            tc_commit_cg2(&empty[stage]);
                            if (kb == kb1 - 1) tc_commit_cg2(&tfull[acc]);
                        }
                        __syncwarp();
                        if (++stage == C2_RING) { stage = 0; phase ^= 1u; }
                    }
                    if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
                }
            }
        }
    } else if (warp >= OZ_EPI_WARP0) {
        // ===== epilogue (both CTAs): warp (4 + 4 h + q) owns TMEM lanes [32 q, 32 q + 32) and tile columns [64 h, 64 h + 64) =====
        const int q = warp & 3, h = (warp - OZ_EPI_WARP0) >> 2;
        TileWalk w; w.init(p);
        int acc = 0; unsigned aphase = 0;
        int tm, tn;
        for (long long idx = pair; w.seek(p, idx, tm, tn); idx += npairs) {
            const int row = tm * C2_BM + (int)rank * OZ_BM + q * 32 + lane;
            const int col0 = tn * OZ_BN + h * 64;
            double accd[MODE == 1 ? 64 : 1];
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 64; ++j) accd[j] = 0.0;
            }
            const int norders = groups + (oz_has_diag(groups) ? 1 : 0);  // the diagonal group is the single product of order `groups`
            for (int t = 0; t < norders; ++t) {
                mbar_wait_bounded(&tfull[acc], aphase);
                tc_fence_after();
                const double sc = __hiloint2double((1023 - OZ_BETA * (t + 2)) << 20, 0);  // 2^(-8 (t+2))
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    unsigned r[32];
                    tc_ld32(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(acc * OZ_BN + h * 64 + c * 32), r);
                    if (MODE == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) accd[c * 32 + j] = fma(exact_i2d((int)r[j]), sc, accd[c * 32 + j]);
                    } else {
                        if (row < p.m) {
                            int* dst = p.Ci + (long long)row * p.ldci + col0 + c * 32;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + c * 32 + j < p.n) dst[j] = (int)r[j];
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));  // the leader's barrier counts both CTAs
                if (++acc == OZ_ACC_STAGES) { acc = 0; aphase ^= 1u; }
            }
            if constexpr (MODE == 1) {
                if (row < p.m) oz_store_row(p, accd, row, col0);
            }
        }
    }

    // neither CTA may exit (or free its TMEM) while the peer can still read its shared memory / signal its barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- variant 4 (default): CTA pairs, ORDER PAIRS SIDE BY SIDE IN ONE N = 256 INSTRUCTION ---------------------------------------
// Variants 1-3 are paced by the operand bytes an instruction pulls out of shared memory (~56 B/clk/SM: M128 x N128 x K32 reads 8 KB
// -> ~145 clk for 64 clk of int8 math; the CTA-pair N = 128 form reads 6 KB -> ~130 clk).  The only lever is more MACs per operand
// byte, i.e. N = 256 -- but a 256-column output tile would need 2 x 128 x 256 fp64 running sums per SM (the whole register file).
// This variant keeps the 256 x 128 output tile of a CTA pair and widens N across the ORDER instead: for the order pair (t, t+1)
//      [ acc_t | acc_{t+1} ] += A_i * [ B_{t-i} ; B_{t+1-i} ]^T ,   i = 0..t,
// one tcgen05.mma.cta_group::2 (M = 256, N = 256) per 32 digits: in a CTA pair the N rows of the B operand are split between the
// two CTAs, so the leader stages the 128 rows of plane t-i and its peer the 128 rows of plane t+1-i of the SAME output columns,
// and both SMs receive the 256-column accumulator [order t | order t+1] for their own 128 rows.  Per SM and instruction that is
// 8 KB of operand reads for 128 clk of math (was 6 KB for 64).  The one product of order t+1 this leaves out, A_{t+1} B_0^T, is an
// N = 128 instruction into the right half (each CTA stages 64 rows of B_0, as in variant 3); an odd plane count runs order 0
// (a single product) the same way, and so does the equal-plane product an even plane count appends (oz_has_diag).  6 planes:
// 9 wide + 4 narrow slot visits per K-block instead of 22 narrow ones.
//   * ring of 6 slots x 32 KB (A 16 KB + B 16 KB; a narrow slot uses 24 KB); every slot is consumed by exactly one instruction
//     group, so there is no "previous slot" coupling as in the paired schedule of variants 1 / 3;
//   * TMEM: 2 buffers x 256 columns [lo | hi]; the epilogue drains buffer b (both orders) while the MMAs fill buffer b^1;
//   * barriers / cluster protocol / tile walk / masks / K-ranges exactly as variant 3.
constexpr int W4_B_BYTES = OZ_BN * OZ_BK;                 // 16 KB: a full 128-row plane tile
constexpr int W4_SLOT_BYTES = C2_A_BYTES + W4_B_BYTES;    // 32 KB
constexpr int W4_RING = 6;
constexpr int W4_ACC = 2;                                 // x 256 TMEM columns
constexpr int W4_SB_STAGE_BYTES = OZ_EPI_WARPS * 64 * (8 + 4);  // per epilogue warp: the 64 column scales of its slab + their exponents
constexpr int W4_SMEM_BYTES = W4_RING * W4_SLOT_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + W4_SB_STAGE_BYTES;
// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 256, M = 256 (cta_group::2)
constexpr unsigned W4_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(256 >> 3) << 17) | ((unsigned)(C2_BM >> 4) << 24);
static_assert(W4_SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

// write-out of one thread's 64 fixed-point sums: exact int64 -> fp64 (two exact halves, ONE rounding), scaled by
// alpha 2^(ea + eb - F - 16), then C (+)= v -- through red.global.add.f64 when p.use_red: every output entry receives exactly one
// add per launch from exactly one thread, so the result is deterministic and no epilogue warp ever waits for a load of C.
// `sbs`: the 64 column scales of the slab, staged in shared memory by the warp (a per-element __ldg of sb cost one exposed L2
// round trip per output entry: 38 % of the MMA warp's time went into waiting for this loop, profiles/r02_ozaki_epilogue.md).
template <int STORE>  // 0: C = v (beta0), 1: red.global.add, 2: load / add / store
__device__ __forceinline__ void oz_store_row_fx_impl(const long long (&accq)[64], double* crow, const double* sbs, double sr,
                                                      int jlo, int jhi) {
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        if (j >= jlo && j < jhi) {
            const int hi = (int)(accq[j] >> 32);
            const unsigned lo = (unsigned)accq[j];
            const double dh = __hiloint2double(0x43300000, (int)((unsigned)hi ^ 0x80000000u)) - 4503601774854144.0;  // exact
            const double dl = __hiloint2double(0x43300000, (int)lo) - 4503599627370496.0;                           // exact
            const double v = fma(dh, 4294967296.0, dl) * (sr * sbs[j]);
            double* dst = crow + j;
            if (STORE == 0) *dst = v;
            else if (STORE == 1) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(dst), "d"(v) : "memory");
            else *dst += v;
        }
    }
}
// Exponent of a scale that is an exact power of two (what the slice kernels produce), or a marker:
//   OZ_E_POISON  : NaN / Inf scale (a poisoned row: the output entry becomes NaN, JAX semantics for a failed factorisation)
//   OZ_E_GENERIC : anything else (zero, negative, denormal, a mantissa) -> the warp takes the FP64 write-out
constexpr int OZ_E_POISON = 0x40000000, OZ_E_GENERIC = 0x40000001;
__device__ __forceinline__ int oz_scale_exp(double s) {
    const int hi = __double2hiint(s), lo = __double2loint(s);
    const int field = (hi >> 20) & 0x7FF;
    if (field == 0x7FF) return OZ_E_POISON;
    if (hi < 0 || field == 0 || ((hi & 0xFFFFF) | lo) != 0) return OZ_E_GENERIC;
    return field - 1023;
}
// INTEGER-ONLY write-out: round-to-nearest-even int64 -> binary64 built by hand, the power-of-two scales applied as an exponent
// add.  Why no FP64 instruction at all: while the tensor pipe runs UTCIMMA the FP64 pipe of the same SM is throttled -- the 5 FP64
// operations per output entry of the FP64 write-out kept the epilogue warps in `stall_math` for a quarter of their time whether the
// kernel issued 17 or 5 of them per entry (profiles/r02_ozaki_epilogue.md section 3), and the MMA warp waited for them.
template <int STORE>  // 0: C = v (beta0), 1: red.global.add, 2: load / add / store
__device__ __forceinline__ void oz_store_row_int_impl(const long long (&accq)[64], double* crow, const int* sbe, int er, int neg,
                                                       int jlo, int jhi) {
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        if (j >= jlo && j < jhi) {
            const long long x = accq[j];
            const unsigned long long a = (unsigned long long)(x < 0 ? -x : x);
            const int ec = sbe[j];
            const int lz = __clzll((long long)a);
            const unsigned long long nrm = a << (lz & 63);          // leading one at bit 63 (a != 0)
            unsigned long long mant = nrm >> 11;                     // 53 bits, implicit one included
            const unsigned rem = (unsigned)nrm & 0x7FFu;
            mant += (rem > 0x400u || (rem == 0x400u && (mant & 1ull))) ? 1ull : 0ull;
            const int E = 63 - lz + 1023 + er + ec;                  // biased exponent of the scaled value
            unsigned long long bits = ((unsigned long long)(unsigned)(E - 1) << 52) + mant;  // a carry out of mant bumps the exponent
            if (a == 0ull || E <= 0) bits = 0ull;                    // zero / below the normal range: flush
            if (E >= 2047) bits = 0x7FF0000000000000ull;
            bits |= (unsigned long long)(unsigned)((x < 0) != (neg != 0)) << 63;
            if (er == OZ_E_POISON || ec == OZ_E_POISON) bits = 0x7FF8000000000000ull;
            const double v = __longlong_as_double((long long)bits);
            double* dst = crow + j;
            if (STORE == 0) *dst = v;
            else if (STORE == 1) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(dst), "d"(v) : "memory");
            else *dst += v;
        }
    }
}
__device__ __forceinline__ void oz_store_row_fx(const OzParams& p, const long long (&accq)[64], int row, int col0, const double* sbs,
                                                const int* sbe, bool generic) {
    const double sr = p.alpha * __ldg(p.sa + row) * __hiloint2double((1023 - p.fx_bits - 2 * OZ_BETA) << 20, 0);
    const long long grow = p.row0 + row;
    double* crow = p.C + (long long)row * p.ldc + col0;  // indexed by the column inside the slab
    bool diag_blk = false;
    if (p.mask == 3) {  // a 64-column slab lies inside one column block (mask_nb is a multiple of 128)
        const long long rb = grow / p.mask_nb, cb = (p.col0 + col0) / p.mask_nb;
        diag_blk = rb == cb;
        if (diag_blk)  // dense diagonal block rb of C2, addressed by (row, column) inside the block
            crow = p.C2 + rb * p.mask_nb * p.mask_nb + (grow - rb * p.mask_nb) * p.mask_nb + (p.col0 + col0 - cb * p.mask_nb);
    }
    // live column range of this row inside the slab (all masks are intervals in the column index)
    int jlo = 0, jhi = p.n - col0 < 64 ? p.n - col0 : 64;
    if (p.mask == 1) {
        const long long last = grow - p.col0 - col0;  // col <= grow - col0
        jhi = last + 1 < jhi ? (int)(last + 1 < 0 ? 0 : last + 1) : jhi;
    } else if ((p.mask == 2 || p.mask == 3) && !diag_blk) {
        const long long first = (grow / p.mask_nb + 1) * p.mask_nb - p.col0 - col0;  // first column of the next block
        jlo = first > 0 ? (first < 64 ? (int)first : 64) : 0;
    }
    const int er = oz_scale_exp(fabs(sr));
    if (generic || er == OZ_E_GENERIC) {  // scales that are not powers of two (never produced by the slice kernels): FP64 write-out
        if (p.beta0) oz_store_row_fx_impl<0>(accq, crow, sbs, sr, jlo, jhi);
        else if (p.use_red) oz_store_row_fx_impl<1>(accq, crow, sbs, sr, jlo, jhi);
        else oz_store_row_fx_impl<2>(accq, crow, sbs, sr, jlo, jhi);
        return;
    }
    const int neg = sr < 0.0;
    if (p.beta0) oz_store_row_int_impl<0>(accq, crow, sbe, er, neg, jlo, jhi);
    else if (p.use_red) oz_store_row_int_impl<1>(accq, crow, sbe, er, neg, jlo, jhi);
    else oz_store_row_int_impl<2>(accq, crow, sbe, er, neg, jlo, jhi);
}

// Order groups of one output tile, enumerated identically by the three roles.  `cursor` runs over the orders t = 0..groups-1:
// a WIDE group covers the order pair (t, t+1) when (groups - t) is even (an odd plane count runs order 0 alone as a single-order
// group).  After the last order an EVEN plane count appends the DIAGONAL group: the one product A_h B_h^T, h = groups / 2, of
// order `groups` (see oz_has_diag).  Every group owns one TMEM buffer.
struct OzGroup {
    int t;        // first order of the group (scale 2^(-8 (t + 2)))
    int wide;     // 1: orders (t, t+1) side by side in N = 256 instructions
    int diag;     // 1: the single product (A_t/2, B_t/2), t == groups
    int nslots;   // ring slots per K-block
    int norders;  // accumulator halves the epilogue drains
};
__device__ __forceinline__ bool oz_next_group(int groups, int& cursor, OzGroup& g) {
    if (cursor < groups) {
        g.t = cursor;
        g.wide = ((groups - cursor) & 1) == 0;
        g.diag = 0;
        g.nslots = g.wide ? cursor + 2 : cursor + 1;
        g.norders = g.wide ? 2 : 1;
        cursor += g.norders;
        return true;
    }
    if (cursor == groups && oz_has_diag(groups)) {
        g.t = groups; g.wide = 0; g.diag = 1; g.nslots = 1; g.norders = 1;
        cursor = groups + 1;
        return true;
    }
    return false;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1)
ozaki_i8_kernel_w4(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzParams p) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);  // identical offset in both CTAs of the pair
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W4_RING * W4_SLOT_BYTES);
    uint64_t* full = bars;                   // [6]   TMA (both CTAs) -> MMA; only the leader's copies are used
    uint64_t* empty = bars + W4_RING;        // [6]   MMA -> TMA, multicast to both CTAs
    uint64_t* tfull = bars + 2 * W4_RING;    // [2]   MMA -> epilogue, multicast to both CTAs
    uint64_t* tempty = tfull + W4_ACC;       // [2]   epilogue (both CTAs) -> MMA; only the leader's copies are used
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tempty + W4_ACC);
    double* sb_stage = reinterpret_cast<double*>(smem + W4_RING * W4_SLOT_BYTES + 256);  // [8 epilogue warps][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < W4_RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < W4_ACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * OZ_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    const int groups = oz_groups(p, MODE);

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own 128 rows of A; wide slot: all 128 rows of ITS plane of B; narrow slot: own 64 rows =====
        TileWalk w; w.init(p);
        int stage = 0; unsigned phase = 0;
        int tm, tn;
        for (long long idx = pair; w.seek(p, idx, tm, tn); idx += npairs) {
            const int m0 = tm * C2_BM + (int)rank * OZ_BM, n0 = tn * OZ_BN;
            int kb0, kb1;
            oz_krange(p, (long long)tm * C2_BM, (long long)tn * OZ_BN, C2_BM, kb0, kb1);
            OzGroup g;
            for (int cursor = 0; oz_next_group(groups, cursor, g);) {
                for (int kb = kb0; kb < kb1; ++kb) {
                    for (int i = 0; i < g.nslots; ++i) {
                        const bool wslot = g.wide && i <= g.t;
                        mbar_wait_bounded(&empty[stage], phase ^ 1u);
                        uint8_t* sA = smem + stage * W4_SLOT_BYTES;
                        if (elect_one()) {
                            if (p.noload) {
                                if (rank == 0) mbar_arrive(&full[stage]);
                            } else {
                                const unsigned fb = mapa_u32(&full[stage], 0);
                                const int xk = kb * OZ_BK;
                                if (wslot) {
                                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * W4_SLOT_BYTES);
                                    const int pb = g.t - i + (int)rank;  // leader: plane t-i (order t), peer: plane t+1-i (order t+1)
                                    tma_load_2d_cg2(sA, &tmA, fb, i * p.pstride + xk, m0);
                                    tma_load_2d_cg2(sA + C2_A_BYTES, &tmB, fb, pb * p.pstride + xk, n0);
                                    tma_load_2d_cg2(sA + C2_A_BYTES + C2_B_BYTES, &tmB, fb, pb * p.pstride + xk, n0 + C2_BNH);
                                } else {
                                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * C2_SLOT_BYTES);
                                    // wide group: the product A_{t+1} B_0 the side-by-side slots leave out; diagonal group: A_h B_h;
                                    // single order: A_i B_{t-i}
                                    const int pa = g.diag ? (g.t >> 1) : (g.wide ? g.t + 1 : i);
                                    const int pb = g.diag ? (g.t >> 1) : (g.wide ? 0 : g.t - i);
                                    tma_load_2d_cg2(sA, &tmA, fb, pa * p.pstride + xk, m0);
                                    tma_load_2d_cg2(sA + C2_A_BYTES, &tmB, fb, pb * p.pstride + xk, n0 + (int)rank * C2_BNH);
                                }
                            }
                        }
                        __syncwarp();
                        if (++stage == W4_RING) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {  // ===== MMA issuer: leader CTA only, one elected lane =====
            TileWalk w; w.init(p);
            int stage = 0; unsigned phase = 0;
            int acc = 0; unsigned aphase = 0;
            int tm, tn;
            for (long long idx = pair; w.seek(p, idx, tm, tn); idx += npairs) {
                int kb0, kb1;
                oz_krange(p, (long long)tm * C2_BM, (long long)tn * OZ_BN, C2_BM, kb0, kb1);
                OzGroup g;
                for (int cursor = 0; oz_next_group(groups, cursor, g);) {
                    mbar_wait_bounded(&tempty[acc], aphase ^ 1u);
                    tc_fence_after();
                    const unsigned d_lo = tmem_base + (unsigned)(acc * 256), d_hi = d_lo + 128u;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        for (int i = 0; i < g.nslots; ++i) {
                            const bool wslot = g.wide && i <= g.t;
                            mbar_wait_bounded(&full[stage], phase);
                            tc_fence_after();
                            const unsigned sA = smem_u32(smem + stage * W4_SLOT_BYTES);
                            if (elect_one()) {
                                const uint64_t da = umma_desc_k_sw128(sA), db = umma_desc_k_sw128(sA + C2_A_BYTES);
                                if (wslot) {
#pragma unroll
                                    for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                        tc_mma_i8_cg2(d_lo, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), W4_IDESC,
                                                      ((kb - kb0) | i | kk) != 0);
                                } else {
                                    // wide group: the right half, already written by slot 0 of this K-block; otherwise the left half,
                                    // first written at (kb0, i = 0)
                                    const unsigned d = g.wide ? d_hi : d_lo;
#pragma unroll
                                    for (int kk = 0; kk < OZ_BK / 32; ++kk)
                                        tc_mma_i8_cg2(d, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), C2_IDESC,
                                                      g.wide ? 1u : (unsigned)(((kb - kb0) | i | kk) != 0));
                                }
                                tc_commit_cg2(&empty[stage]);
                                if (kb == kb1 - 1 && i == g.nslots - 1) tc_commit_cg2(&tfull[acc]);
                            }
                            __syncwarp();
                            if (++stage == W4_RING) { stage = 0; phase ^= 1u; }
                        }
                    }
                    if (++acc == W4_ACC) { acc = 0; aphase ^= 1u; }
                }
            }
        }
    } else if (warp >= OZ_EPI_WARP0) {
        // ===== epilogue (both CTAs): warp (4 + 4 h + q) owns TMEM lanes [32 q, 32 q + 32) and tile columns [64 h, 64 h + 64) =====
        const int q = warp & 3, h = (warp - OZ_EPI_WARP0) >> 2;
        TileWalk w; w.init(p);
        int acc = 0; unsigned aphase = 0;
        int tm, tn;
        for (long long idx = pair; w.seek(p, idx, tm, tn); idx += npairs) {
            const int row = tm * C2_BM + (int)rank * OZ_BM + q * 32 + lane;
            const int col0 = tn * OZ_BN + h * 64;
            long long accq[MODE == 1 ? 64 : 1];  // exact fixed-point running sums (see oz_fx_bits)
            double* sbs = sb_stage + (warp - OZ_EPI_WARP0) * 64;
            int* sbe = reinterpret_cast<int*>(sb_stage + OZ_EPI_WARPS * 64) + (warp - OZ_EPI_WARP0) * 64;
            bool generic = false;  // warp-uniform: some column scale of the slab is not a power of two
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 64; ++j) accq[j] = 0;
                __syncwarp();  // the previous tile's write-out has read its scales
                const double s0 = col0 + lane < p.n ? __ldg(p.sb + col0 + lane) : 1.0;
                const double s1 = col0 + 32 + lane < p.n ? __ldg(p.sb + col0 + 32 + lane) : 1.0;
                const int e0 = oz_scale_exp(s0), e1 = oz_scale_exp(s1);
                sbs[lane] = s0; sbs[32 + lane] = s1;
                sbe[lane] = e0; sbe[32 + lane] = e1;
                generic = __any_sync(0xffffffffu, e0 == OZ_E_GENERIC || e1 == OZ_E_GENERIC) != 0;
                __syncwarp();
            }
            OzGroup g;
            for (int cursor = 0; oz_next_group(groups, cursor, g);) {
                mbar_wait_bounded(&tfull[acc], aphase);
                tc_fence_after();
                for (int o = 0; o < g.norders; ++o) {
                    const int sh = p.fx_bits - OZ_BETA * (g.t + o);  // weight 2^sh of this order (warp-uniform)
                    const unsigned tcol = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(acc * 256 + o * 128 + h * 64);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        unsigned r[16];
                        tc_ld16(tcol + (unsigned)(c * 16), r);
                        if (MODE == 1) {
                            if (sh >= 32) {  // order 0: lands in the high word
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    accq[c * 16 + j] += (long long)((unsigned long long)(unsigned)((int)r[j] << (sh - 32)) << 32);
                            } else if (sh >= 0) {  // one IMAD.WIDE per value
                                const int mul = 1 << sh;
#pragma unroll
                                for (int j = 0; j < 16; ++j) accq[c * 16 + j] += (long long)(int)r[j] * (long long)mul;
                            } else {  // below the unit: round to nearest
                                const int rs = -sh, half = 1 << (rs - 1);
#pragma unroll
                                for (int j = 0; j < 16; ++j) accq[c * 16 + j] += (long long)(((int)r[j] + half) >> rs);
                            }
                        } else {
                            if (row < p.m) {
                                int* dst = p.Ci + (long long)row * p.ldci + col0 + c * 16;
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (col0 + c * 16 + j < p.n) dst[j] = (int)r[j];
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(&tempty[acc], 0));  // the leader's barrier counts both CTAs
                if (++acc == W4_ACC) { acc = 0; aphase ^= 1u; }
            }
            if constexpr (MODE == 1) {
                if (row < p.m) oz_store_row_fx(p, accq, row, col0, sbs, sbe, generic);
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- digit extraction ----------------------------------------------------------------------------------------------
// Balanced radix-256 digits of R = x 2^-e (|R| <= OZ_RMAX): ONE rounding to the last plane, I = rint(R 2^(8 s)) (|I| < 2^55 for
// s <= 7, exact in int64), then the digits are peeled from the least significant end with a carry,
//      d = ((I + 128) mod 256) - 128  in [-128, 127],      I <- (I - d) / 256,
// so every plane uses the whole int8 range and sum_p d_p 256^(s-1-p) == I exactly.  The top digit stays inside int8 because
// the row exponent leaves |R| <= 0.494 < (127 - 128/255) / 256.  Truncating the planes after the first s' < s leaves a remainder of
// at most 0.502 units of plane s' (the device-side plane guard uses a prefix of the extracted planes).
// one CTA per row: exponent from the row maximum, then the digits of every entry
__global__ void __launch_bounds__(128) ozaki_slice_kernel(long long rows, int k, int kplane, const double* __restrict__ X,
                                                          long long ldx, int nslices, signed char* __restrict__ Q,
                                                          long long ldq, double* __restrict__ scale,
                                                          const int* __restrict__ nslices_dev) {
    __shared__ double red[4];
    const long long r = blockIdx.x;
    if (r >= rows) return;
    if (nslices_dev) {  // the plane count the product will use: round THERE (uniform across the grid)
        const int v = __ldg(nslices_dev);
        nslices = (v < 1 || v > nslices) ? nslices : v;
    }
    const double* x = X + r * ldx;
    double mx = 0.0;
    bool bad = false;
    for (int c = threadIdx.x; c < k; c += 128) {
        double v = fabs(x[c]);
        bad |= !(v <= 1.7976931348623157e308);  // NaN or Inf
        mx = fmax(mx, v);
    }
    mx = bad ? __longlong_as_double(0x7ff8000000000000LL) : mx;
    // NaN-propagating max across the CTA
    for (int o = 16; o > 0; o >>= 1) {
        double other = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = (mx != mx || other != other) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(mx, other);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 4; ++i) mx = (mx != mx || red[i] != red[i]) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(mx, red[i]);
    int e = 0;
    double sc;
    if (mx != mx) {
        sc = mx;  // NaN row scale: the product becomes NaN (JAX semantics for a failed factorisation)
    } else if (mx == 0.0) {
        sc = 1.0;
    } else {
        e = oz_row_exponent(mx);  // |x| 2^-e <= 0.494
        sc = scalbn(1.0, e);
    }
    if (threadIdx.x == 0) scale[r] = sc;
    signed char* qrow = Q + r * ldq;
    for (int c = threadIdx.x; c < kplane; c += 128) {
        const double R = (mx != mx || c >= k) ? 0.0 : scalbn(x[c], -e);  // columns [k, kplane) are zero padding
        long long I = oz_fixed_point(R, nslices);
        for (int p = nslices - 1; p >= 0; --p) {
            const int d = (int)((I + 128) & 255) - 128;
            qrow[(long long)p * kplane + c] = (signed char)d;
            I = (I - d) >> 8;
        }
    }
}

// ---- digit extraction along COLUMNS (contraction over the rows of X): Qt[c, p*kplane + r] = digit p of X[r, c] ------------
// used by the SGPR statistics SYRK  K_b^T K_b  (operand "rows" are the columns of the rows x M block, K = block rows)
__global__ void __launch_bounds__(256) col_absmax_kernel(long long rows, int cols, const double* __restrict__ X, long long ldx,
                                                         unsigned long long* __restrict__ colmax_bits) {
    // |x| as an unsigned 64-bit pattern is order preserving for non-negative doubles, Inf < NaN patterns: the max propagates NaN
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const long long r_begin = (long long)blockIdx.y * 1024;
    const long long r_end = r_begin + 1024 < rows ? r_begin + 1024 : rows;
    unsigned long long m = 0ull;
    if (c < cols)
        for (long long r = r_begin + (threadIdx.x >> 5); r < r_end; r += 8) {
            unsigned long long b = (unsigned long long)__double_as_longlong(fabs(X[r * ldx + c]));
            m = b > m ? b : m;
        }
    __shared__ unsigned long long red[8][32];
    red[threadIdx.x >> 5][threadIdx.x & 31] = m;
    __syncthreads();
    if (threadIdx.x < 32 && c < cols) {
        for (int i = 1; i < 8; ++i) m = red[i][threadIdx.x] > m ? red[i][threadIdx.x] : m;
        atomicMax(colmax_bits + c, m);
    }
}
// tile of 128 rows x 32 columns per CTA: coalesced reads along the columns, digits staged in shared memory, 128-byte
// contiguous writes along r for every (plane, column); rows >= `rows` (up to kplane) are written as zero digits
__global__ void __launch_bounds__(256) ozaki_slice_t_kernel(long long rows, int cols, long long kplane, const double* __restrict__ X,
                                                            long long ldx, int nslices, signed char* __restrict__ Qt, long long ldq,
                                                            const unsigned long long* __restrict__ colmax_bits,
                                                            double* __restrict__ scale) {
    __shared__ signed char dig[OZ_PLANES_MAX][32][128 + 16];
    const int c0 = blockIdx.x * 32;
    const long long r0 = (long long)blockIdx.y * 128;
    const int lc = threadIdx.x & 31;
    const int c = c0 + lc;
    double mx = 0.0;
    int e = 0;
    bool nanrow = false;
    if (c < cols) {
        mx = __longlong_as_double((long long)colmax_bits[c]);
        nanrow = !(mx <= 1.7976931348623157e308);
        if (!nanrow && mx != 0.0) e = oz_row_exponent(mx);
        if (blockIdx.y == 0 && threadIdx.x < 32)
            scale[c] = nanrow ? __longlong_as_double(0x7ff8000000000000LL) : (mx == 0.0 ? 1.0 : scalbn(1.0, e));
    }
    for (int lr = threadIdx.x >> 5; lr < 128; lr += 8) {
        const long long r = r0 + lr;
        const double R = (c < cols && r < rows && !nanrow) ? scalbn(X[r * ldx + c], -e) : 0.0;
        long long I = oz_fixed_point(R, nslices);
        for (int p = nslices - 1; p >= 0; --p) {
            const int d = (int)((I + 128) & 255) - 128;
            dig[p][lc][lr] = (signed char)d;
            I = (I - d) >> 8;
        }
    }
    __syncthreads();
    // write-out: (plane, column) rows of 128 bytes, 4 bytes per thread
    const int total = nslices * 32 * 32;  // 4-byte words
    for (int w = threadIdx.x; w < total; w += 256) {
        const int p = w / (32 * 32), rem = w % (32 * 32), cc = rem / 32, word = rem % 32;
        if (c0 + cc < cols)
            *reinterpret_cast<int*>(Qt + (long long)(c0 + cc) * ldq + (long long)p * kplane + r0 + word * 4) =
                *reinterpret_cast<const int*>(&dig[p][cc][word * 4]);
    }
}

// ---- out_w[c] += sum_r w[r] X[r,c],  out_1[c] += sum_r X[r,c]  (the two augmented rows of the SGPR statistics) -------------
// two stages so the summation order is fixed: per-1024-row partials, then an ordered reduction
__global__ void __launch_bounds__(256) col_wsum_partial_kernel(long long rows, int cols, const double* __restrict__ X, long long ldx,
                                                               const double* __restrict__ w, double* __restrict__ part) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const long long r_begin = (long long)blockIdx.y * 1024;
    const long long r_end = r_begin + 1024 < rows ? r_begin + 1024 : rows;
    double sw = 0.0, s1 = 0.0;
    if (c < cols)
        for (long long r = r_begin + (threadIdx.x >> 5); r < r_end; r += 8) {
            const double x = X[r * ldx + c];
            sw = fma(w[r], x, sw);
            s1 += x;
        }
    __shared__ double red[2][8][32];
    red[0][threadIdx.x >> 5][threadIdx.x & 31] = sw;
    red[1][threadIdx.x >> 5][threadIdx.x & 31] = s1;
    __syncthreads();
    if (threadIdx.x < 32 && c < cols) {
        for (int i = 1; i < 8; ++i) { sw += red[0][i][threadIdx.x]; s1 += red[1][i][threadIdx.x]; }
        part[((long long)blockIdx.y * 2 + 0) * cols + c] = sw;
        part[((long long)blockIdx.y * 2 + 1) * cols + c] = s1;
    }
}
__global__ void col_wsum_reduce_kernel(int chunks, int cols, const double* __restrict__ part, double* __restrict__ out_w,
                                       double* __restrict__ out_1) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double sw = 0.0, s1 = 0.0;
    for (int k = 0; k < chunks; ++k) { sw += part[((long long)k * 2 + 0) * cols + c]; s1 += part[((long long)k * 2 + 1) * cols + c]; }
    out_w[c] += sw;
    out_1[c] += s1;
}

// same reduction for many chunks (the fused Gram -> digits kernel leaves one partial per 64-row tile row: 1024 per block): eight
// row groups walk interleaved chunk subsets, then a fixed-order sum of the eight -- deterministic, 8x the parallelism
__global__ void __launch_bounds__(256) col_partials_reduce_kernel(int chunks, int cols, const double* __restrict__ part,
                                                                  double* __restrict__ out_w, double* __restrict__ out_1) {
    __shared__ double red[2][8][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double sw = 0.0, s1 = 0.0;
    if (c < cols)
        for (int k = threadIdx.y; k < chunks; k += 8) { sw += part[((long long)k * 2 + 0) * cols + c]; s1 += part[((long long)k * 2 + 1) * cols + c]; }
    red[0][threadIdx.y][threadIdx.x] = sw;
    red[1][threadIdx.y][threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int i = 1; i < 8; ++i) { sw += red[0][i][threadIdx.x]; s1 += red[1][i][threadIdx.x]; }
        out_w[c] += sw;
        out_1[c] += s1;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        if (f && q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
// digit matrix [rows, width] int8, row stride ld bytes -> 2-D map with a {128 B, 128 rows} box
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t width, int64_t ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return GPB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 15)) return GPB_ERR_INVALID;
    cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GPB_OK : GPB_ERR_INVALID;
}
int sm_count() {  // per device: one process may drive several GPUs (XLA's per-device threads)
    static int n[GPB_MAX_DEVICES] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}
// kernel variant: 4 = CTA pairs with order pairs side by side in N = 256 instructions (default), 3 = CTA pairs / N = 128 (round-2
// first half), 1 = one CTA per tile (round-1 kernel), 2 = plane-resident 128 x 64 tiles.  GPB_OZ_KERNEL selects 1 / 2 / 3 for
// measurements (profiles/r02_ozaki_cg2.md, r02_ozaki_w4.md); read once, the value never changes afterwards.
int variant() {
    static const int v = [] {
        const char* e = std::getenv("GPB_OZ_KERNEL");
        const int x = e ? std::atoi(e) : 4;
        return (x >= 1 && x <= 4) ? x : 4;
    }();
    return v;
}
template <int MODE>
int launch(stream_t s, const void* A, int64_t rowsA, int64_t lda, const void* B, int64_t rowsB, int64_t ldb, int64_t width,
           OzParams& p) {
    int v = variant();
    if (v == 4 && std::getenv("GPB_OZ_NOLOAD")) v = 3;  // the no-load pacing hook exists in variants 1-3 only
    // cudaFuncSetAttribute is per device: remember it per (device, variant), never per process
    static bool attr_set[GPB_MAX_DEVICES][5] = {};
    const int dev = current_device();
    if (!attr_set[dev][v]) {
        cudaError_t e = v == 1   ? cudaFuncSetAttribute(ozaki_i8_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES)
                        : v == 2 ? cudaFuncSetAttribute(ozaki_i8_kernel_v2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, O2_SMEM_BYTES)
                        : v == 3 ? cudaFuncSetAttribute(ozaki_i8_kernel_cg2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES)
                                 : cudaFuncSetAttribute(ozaki_i8_kernel_w4<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, W4_SMEM_BYTES);
        if (e != cudaSuccess) return GPB_ERR_LAUNCH;
        attr_set[dev][v] = true;
    }
    p.bn = v == 2 ? O2_BN : OZ_BN;
    p.bm = v >= 3 ? C2_BM : OZ_BM;
    p.kbs = (p.kblocks % OZ_KBS_MAX == 0 && !std::getenv("GPB_OZ_KBS1")) ? OZ_KBS_MAX : 1;
    {
        static int pair = [] { const char* e = std::getenv("GPB_OZ_PAIR"); return (e && std::atoi(e) == 0) ? 0 : 1; }();
        p.pair = pair;
    }
    if (p.pair && p.nslices > 1) {  // paired orders: ring slots of one (A, B) K-block pair
        p.kbs = 1; p.nstages = OZ_MAX_RING; p.stage_bytes = OZ_KB_BYTES;
    } else {
        p.pair = 0; p.nstages = OZ_STAGES; p.stage_bytes = OZ_STAGE_BYTES;
    }
    {
        static int noload = [] { const char* e = std::getenv("GPB_OZ_NOLOAD"); return (e && std::atoi(e) == 1) ? 1 : 0; }();
        p.noload = noload;
        static int red = [] { const char* e = std::getenv("GPB_OZ_RED"); return (e && std::atoi(e) == 0) ? 0 : 1; }();
        p.use_red = red;
        p.fx_bits = oz_fx_bits((long long)p.kblocks * OZ_BK);
    }
    CUtensorMap ta, tb;
    int rc = make_map(&ta, A, rowsA, width, lda, OZ_BM);
    if (rc) return rc;
    rc = make_map(&tb, B, rowsB, width, ldb, v >= 3 ? C2_BNH : p.bn);
    if (rc) return rc;
    p.ntm = (p.m + p.bm - 1) / p.bm;
    p.ntn = (p.n + p.bn - 1) / p.bn;
    long long tiles = (long long)p.ntm * p.ntn;  // upper bound on the live tiles; CTAs beyond the live count exit at once
    const int ctas_per_tile = v >= 3 ? 2 : 1;
    int units = sm_count() / ctas_per_tile;  // v3 / v4: one CTA pair per TPC
    if (p.max_ctas > 0 && p.max_ctas / ctas_per_tile < units) units = p.max_ctas / ctas_per_tile > 0 ? p.max_ctas / ctas_per_tile : 1;
    int grid = (int)(tiles < units ? tiles : units) * ctas_per_tile;
    if (grid <= 0) return GPB_OK;
    const bool prof = profile_enabled();
    if (prof) {  // algorithmic int8 operations: live output entries x K x digit pairs x 2
        double live = 0.0;
        for (long long i = 0; i < p.m; ++i) {
            long long c = p.n;
            if (p.mask == 1) {
                c = p.row0 + i - p.col0 + 1;
                c = c < 0 ? 0 : (c > p.n ? p.n : c);
            } else if (p.mask == 2 || p.mask == 3) {
                long long first = ((p.row0 + i) / p.mask_nb + (p.mask == 2 ? 1 : 0)) * p.mask_nb - p.col0;
                first = first < 0 ? 0 : (first > p.n ? p.n : first);
                c = p.n - first;
            }
            live += (double)c;
        }
        const int S = MODE == 0 ? 1 : p.nslices;
        double kfrac = 1.0;  // triangular K-range: fraction of the K-blocks an average tile visits
        if (p.krange != KR_FULL && p.kblocks > 0) {
            const bool on_b = p.krange == KR_B_LOWER || p.krange == KR_B_UPPER;
            const int nt = on_b ? p.ntn : p.ntm, step = on_b ? OZ_BN : p.bm;
            double acc = 0.0;
            for (int t = 0; t < nt; ++t) {
                long long lo = 0, hi = p.kblocks;
                const long long x0 = (long long)t * step;
                if (p.krange == KR_B_LOWER || p.krange == KR_A_LOWER) hi = (x0 + step - 1 + p.kr_off) / OZ_BK + 1;
                else lo = (x0 + p.kr_off) / OZ_BK;
                lo = lo < 0 ? 0 : (lo > p.kblocks - 1 ? p.kblocks - 1 : lo);
                hi = hi < 1 ? 1 : (hi > p.kblocks ? p.kblocks : hi);
                acc += (double)(hi - lo);
            }
            kfrac = acc / ((double)nt * p.kblocks);
        }
        profile_ozaki_begin(s, 2.0 * live * kfrac * (double)p.kblocks * OZ_BK * (S * (S + 1) / 2 + (v >= 3 && oz_has_diag(S) ? 1 : 0)));
    }
    if (v == 1) ozaki_i8_kernel<MODE><<<grid, OZ_THREADS, OZ_SMEM_BYTES, to_stream(s)>>>(ta, tb, p);
    else if (v == 2) ozaki_i8_kernel_v2<MODE><<<grid, OZ_THREADS, O2_SMEM_BYTES, to_stream(s)>>>(ta, tb, p);
    else if (v == 3) ozaki_i8_kernel_cg2<MODE><<<grid, OZ_THREADS, C2_SMEM_BYTES, to_stream(s)>>>(ta, tb, p);
    else ozaki_i8_kernel_w4<MODE><<<grid, OZ_THREADS, W4_SMEM_BYTES, to_stream(s)>>>(ta, tb, p);
    if (prof) profile_gemm_end(s);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace

int ozaki_slice(stream_t s, int64_t rows, int64_t k, int64_t kplane, const double* X, int64_t ldx, int nslices, int8_t* Q,
                int64_t ldq, double* scale, const int* nslices_dev) {
    if (rows < 0 || k <= 0 || kplane < k || nslices < 1 || nslices > OZ_PLANES_MAX || !X || !Q || !scale || ldq < (int64_t)nslices * kplane)
        return GPB_ERR_INVALID;
    if (rows == 0) return GPB_OK;
    ozaki_slice_kernel<<<(unsigned)rows, 128, 0, to_stream(s)>>>(rows, (int)k, (int)kplane, X, ldx, nslices,
                                                                 reinterpret_cast<signed char*>(Q), ldq, scale, nslices_dev);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int ozaki_slice_t(stream_t s, int64_t rows, int64_t cols, int64_t kplane, const double* X, int64_t ldx, int nslices, int8_t* Qt,
                  int64_t ldq, double* scale, double* colmax_scratch) {
    if (rows <= 0 || cols <= 0 || kplane < rows || kplane % 128 || nslices < 1 || nslices > OZ_PLANES_MAX || !X || !Qt || !scale ||
        !colmax_scratch || ldq < (int64_t)nslices * kplane || (ldq & 3) || (reinterpret_cast<uintptr_t>(Qt) & 3))
        return GPB_ERR_INVALID;
    if (cudaMemsetAsync(colmax_scratch, 0, (size_t)cols * sizeof(double), to_stream(s)) != cudaSuccess) return GPB_ERR_LAUNCH;
    dim3 g1((unsigned)((cols + 31) / 32), (unsigned)((rows + 1023) / 1024));
    col_absmax_kernel<<<g1, 256, 0, to_stream(s)>>>(rows, (int)cols, X, ldx, reinterpret_cast<unsigned long long*>(colmax_scratch));
    GPB_LAUNCH_CHECK();
    dim3 g2((unsigned)((cols + 31) / 32), (unsigned)(kplane / 128));
    ozaki_slice_t_kernel<<<g2, 256, 0, to_stream(s)>>>(rows, (int)cols, kplane, X, ldx, nslices, reinterpret_cast<signed char*>(Qt),
                                                        ldq, reinterpret_cast<const unsigned long long*>(colmax_scratch), scale);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int col_weighted_sums(stream_t s, int64_t rows, int64_t cols, const double* X, int64_t ldx, const double* w, double* scratch,
                      double* out_w, double* out_1) {
    if (rows <= 0 || cols <= 0 || !X || !w || !scratch || !out_w || !out_1) return GPB_ERR_INVALID;
    const int chunks = (int)((rows + 1023) / 1024);
    dim3 g((unsigned)((cols + 31) / 32), (unsigned)chunks);
    col_wsum_partial_kernel<<<g, 256, 0, to_stream(s)>>>(rows, (int)cols, X, ldx, w, scratch);
    GPB_LAUNCH_CHECK();
    col_wsum_reduce_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, to_stream(s)>>>(chunks, (int)cols, scratch, out_w, out_1);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int col_partials_reduce(stream_t s, int64_t chunks, int64_t cols, const double* part, double* out_w, double* out_1) {
    if (chunks <= 0 || cols <= 0 || !part || !out_w || !out_1) return GPB_ERR_INVALID;
    col_partials_reduce_kernel<<<(unsigned)((cols + 31) / 32), dim3(32, 8), 0, to_stream(s)>>>((int)chunks, (int)cols, part, out_w, out_1);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int igemm_i8(stream_t s, int64_t m, int64_t n, int64_t k, const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* C,
             int64_t ldc) {
    if (m < 0 || n < 0 || k <= 0 || !A || !B || !C) return GPB_ERR_INVALID;
    if (k % OZ_BK) return GPB_ERR_UNSUPPORTED;
    if (m == 0 || n == 0) return GPB_OK;
    OzParams p = {};
    p.m = (int)m; p.n = (int)n; p.kblocks = (int)(k / OZ_BK); p.nslices = 1;
    p.Ci = C; p.ldci = ldc; p.pstride = (int)k;
    return launch<0>(s, A, m, lda, B, n, ldb, k, p);
}

int ozaki_gemm(stream_t s, const OzakiGemmDesc& d) {
    if (d.M < 0 || d.N < 0 || d.K <= 0 || d.nslices < 1 || d.nslices > OZ_PLANES_MAX || !d.Qa || !d.Qb || !d.sa || !d.sb || !d.C)
        return GPB_ERR_INVALID;
    // int32 accumulator headroom of the deepest order: nslices * K * 128^2 must stay below 2^31
    if (d.K % OZ_BK || (int64_t)d.nslices * d.K * OZ_DIGIT_SQ_MAX >= (1ll << 31)) return GPB_ERR_UNSUPPORTED;
    if (d.mask != MASK_NONE && d.mask != MASK_LOWER && d.mask != MASK_BLOCK_STRICT_UPPER && d.mask != MASK_BLOCK_UPPER_DIAG_TO_C2)
        return GPB_ERR_UNSUPPORTED;
    const bool ext = d.krange != KR_FULL || d.mask == MASK_BLOCK_UPPER_DIAG_TO_C2;
    if (ext && !ozaki_supports_extensions()) return GPB_ERR_UNSUPPORTED;
    if (d.mask == MASK_BLOCK_UPPER_DIAG_TO_C2 && (!d.C2 || d.mask_nb <= 0 || d.mask_nb % 128 || d.mask_col0 % 128)) return GPB_ERR_INVALID;
    if (d.krange < KR_FULL || d.krange > KR_A_UPPER) return GPB_ERR_INVALID;
    if (d.M == 0 || d.N == 0) return GPB_OK;
    OzParams p = {};
    p.m = (int)d.M; p.n = (int)d.N; p.kblocks = (int)(d.K / OZ_BK); p.nslices = d.nslices;
    p.mask = d.mask == MASK_LOWER ? 1 : (d.mask == MASK_BLOCK_STRICT_UPPER ? 2 : (d.mask == MASK_BLOCK_UPPER_DIAG_TO_C2 ? 3 : 0));
    p.row0 = d.mask_row0; p.col0 = d.mask_col0; p.mask_nb = d.mask_nb > 0 ? d.mask_nb : 1;
    p.C2 = d.C2; p.krange = d.krange; p.kr_off = d.kr_off; p.max_ctas = d.max_ctas;
    p.C = d.C; p.ldc = d.ldc; p.sa = d.sa; p.sb = d.sb; p.alpha = d.alpha; p.beta0 = d.beta0; p.planes_dev = d.nslices_dev;
    const int64_t ps = d.plane_stride > 0 ? d.plane_stride : d.K;
    if (ps < d.K || ps % 16) return GPB_ERR_INVALID;
    p.pstride = (int)ps;
    return launch<1>(s, d.Qa, d.M, d.ldqa, d.Qb, d.N, d.ldqb, (int64_t)(d.nslices - 1) * ps + d.K, p);
}

bool ozaki_available() { return encode_tiled() != nullptr; }
int device_sm_count() { return sm_count(); }
bool ozaki_supports_extensions() { return variant() >= 3; }

namespace {
__global__ void ozaki_choose_planes_kernel(int requested, double n, const double* __restrict__ variance,
                                           const double* __restrict__ obs_stddev, double jitter, int* __restrict__ out) {
    int planes = requested;
    if (requested == OZ_AUTO) {
        planes = OZ_AUTO_PLANES_HI;
        if (variance && obs_stddev) {
            const double s = obs_stddev[0] * obs_stddev[0] + jitter;
            const double bound = (n * fabs(variance[0]) + s) / s;  // NaN / s == 0 compare false -> the larger plane count
            if (bound <= OZ_AUTO_COND_LIMIT) planes = OZ_AUTO_PLANES_LO;
        }
    }
    out[0] = planes;
}
}  // namespace

int ozaki_choose_planes(stream_t s, int requested, int64_t N, const double* variance, const double* obs_stddev, double jitter,
                        int* planes_out) {
    if (!planes_out || N < 0 || !(requested == OZ_AUTO || (requested >= 1 && requested <= OZ_PLANES_MAX))) return GPB_ERR_INVALID;
    ozaki_choose_planes_kernel<<<1, 1, 0, to_stream(s)>>>(requested, (double)N, variance, obs_stddev, jitter, planes_out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int ozaki_auto_planes_host(int64_t N, double variance, double obs_stddev, double jitter) {
    const double s = obs_stddev * obs_stddev + jitter;
    const double bound = ((double)N * (variance < 0 ? -variance : variance) + s) / s;
    return bound <= OZ_AUTO_COND_LIMIT ? OZ_AUTO_PLANES_LO : OZ_AUTO_PLANES_HI;
}

}  // namespace gpb
