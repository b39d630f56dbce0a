// Single-CTA Cholesky + triangular inverse of one diagonal leaf block (n <= 128) held entirely in
// shared memory.  This is the latency-bound POTRF "panel" step of the blocked right-looking
// factorisation (reference: jnp.linalg.cholesky, gpjax/linalg/operations.py:55).  The explicit
// inverse it also emits turns every TRSM of the blocked algorithms into a DMMA GEMM.
//
// Both phases are blocked in 8-column panels so the number of CTA-wide barriers is O(n/8), not O(n):
//   * factor : every row thread factors the 8x8 diagonal block REDUNDANTLY in registers (no
//              barrier inside a panel), forward-substitutes its own row, then all threads apply the
//              rank-8 update to the trailing lower triangle;
//   * inverse: right-to-left panels, X21 = -X22 L21 inv(L11) with the 8x8 inverse again formed
//              redundantly in registers and the X22*L21 product split over two threads per row.
//
// Failure semantics mirror JAX: a non-positive or NaN pivot NaN-fills the outputs; *info records
// the (1-based, global) index of the first failing pivot.
#include "common.cuh"

namespace gpb {

namespace {

constexpr int LEAF = 128;
constexpr int LDSM = LEAF + 1;  // odd stride: column walks are bank-conflict free
constexpr int LT = 256;         // threads: 2 per row
constexpr int PW = 8;           // panel width

// Cholesky of the PW x PW block D (lower part) in registers.  Returns a bitmask of failed pivots.
__device__ __forceinline__ unsigned chol8(const double (&D)[PW][PW], double (&L)[PW][PW], double (&inv)[PW]) {
    unsigned bad = 0;
#pragma unroll
    for (int c = 0; c < PW; ++c) {
        double s = D[c][c];
#pragma unroll
        for (int k = 0; k < c; ++k) s = fma(-L[c][k], L[c][k], s);
        if (!(s > 0.0)) bad |= (1u << c);
        inv[c] = rsqrt(s);   // one dependent special-function step per pivot instead of sqrt + divide
        L[c][c] = s * inv[c];
#pragma unroll
        for (int r = c + 1; r < PW; ++r) {
            double t = D[r][c];
#pragma unroll
            for (int k = 0; k < c; ++k) t = fma(-L[r][k], L[c][k], t);
            L[r][c] = t * inv[c];
        }
    }
    return bad;
}

// Inverse of a lower-triangular PW x PW block in registers.
__device__ __forceinline__ void trinv8(const double (&L)[PW][PW], double (&X)[PW][PW]) {
    double rinv[PW];
#pragma unroll
    for (int c = 0; c < PW; ++c) rinv[c] = 1.0 / L[c][c];  // independent reciprocals, off the dependency chain
#pragma unroll
    for (int c = 0; c < PW; ++c) {
        X[c][c] = rinv[c];
#pragma unroll
        for (int r = c + 1; r < PW; ++r) {
            double t = 0.0;
#pragma unroll
            for (int k = c; k < r; ++k) t = fma(L[r][k], X[k][c], t);
            X[r][c] = -t * rinv[r];
        }
    }
}

__global__ void __launch_bounds__(LT, 1) potrf_leaf_kernel(int n, double* __restrict__ A, int64_t lda,
                                                           double* __restrict__ Dinv, int64_t ldd,
                                                           double* __restrict__ DinvT, int64_t lddt, int* info,
                                                           int64_t global_row0, int factor) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;  // [LEAF][LDSM]; only the lower triangle is ever read
    const int tid = threadIdx.x;
    const int i = tid >> 1;  // row owned by this thread: the two threads of a row are adjacent lanes,
    const int q = tid & 1;   // so their partial sums combine with one shuffle instead of shared memory

    // lower triangle of the block, identity padding beyond n (keeps partial panels well defined)
    {
        const int c = tid & (LEAF - 1), r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < LEAF; r += 2) {  // independent, coalesced loads: 8 in flight per thread
            double v = 0.0;
            if (r < n && c <= r) v = A[(int64_t)r * lda + c];
            else if (r >= n && c == r) v = 1.0;
            S[r * LDSM + c] = v;
        }
    }
    __syncthreads();

    int fail = 0;
    if (factor) {
        for (int p0 = 0; p0 < n; p0 += PW) {
            unsigned bad = 0;
            double l[PW];
            const bool active = (q == 0 && i >= p0 && i < n);
            if (active) {
                double D[PW][PW], L[PW][PW], inv[PW];
#pragma unroll
                for (int r = 0; r < PW; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) D[r][c] = S[(p0 + r) * LDSM + p0 + c];
                bad = chol8(D, L, inv);
#pragma unroll
                for (int c = 0; c < PW; ++c) {
                    double t = S[i * LDSM + p0 + c];
#pragma unroll
                    for (int k = 0; k < c; ++k) t = fma(-l[k], L[c][k], t);
                    t *= inv[c];
                    if (i == p0 + c) t = L[c][c];  // exact diagonal
                    if (i < p0 + c) t = 0.0;       // strict upper part of the diagonal block
                    l[c] = t;
                }
            }
            __syncthreads();  // every thread has read the un-factored diagonal block before its rows change
            if (active) {
#pragma unroll
                for (int c = 0; c < PW; ++c) S[i * LDSM + p0 + c] = l[c];
            }
            const int anybad = __syncthreads_or((int)bad);  // also publishes the panel
            if (anybad) {
                fail = p0 + __ffs(anybad);  // 1-based index of the first failing pivot
                break;
            }
            // rank-PW update of the trailing lower triangle
            if (i >= p0 + PW && i < n) {
#pragma unroll
                for (int k = 0; k < PW; ++k) l[k] = S[i * LDSM + p0 + k];
                for (int c = p0 + PW + q; c <= i; c += 2) {
                    double t = S[i * LDSM + c];
                    const double* lc = S + c * LDSM + p0;  // panel entries of row c
#pragma unroll
                    for (int k = 0; k < PW; ++k) t = fma(-l[k], lc[k], t);
                    S[i * LDSM + c] = t;
                }
            }
            __syncthreads();
        }
    }

    if (fail != 0) {
        const double qnan = nan("");
        if (tid == 0 && info && *info == 0) *info = (int)(global_row0 + fail);
        for (int idx = tid; idx < n * n; idx += LT) {
            int r = idx / n, c = idx % n;
            if (c <= r) A[(int64_t)r * lda + c] = qnan;
            if (Dinv) Dinv[(int64_t)r * ldd + c] = qnan;
            if (DinvT) DinvT[(int64_t)r * lddt + c] = qnan;
        }
        return;
    }

    // ---- write L back (lower triangle only) --------------------------------------------------
    if (factor) {
        const int c = tid & (LEAF - 1), r0 = tid >> 7;
#pragma unroll 8
        for (int r = r0; r < LEAF; r += 2)
            if (r < n && c <= r) A[(int64_t)r * lda + c] = S[r * LDSM + c];
    }
    if (!Dinv && !DinvT) return;
    __syncthreads();

    // ---- in-place inverse, panels from the right ------------------------------------------------
    // X21 = -(X22 L21) inv(L11).  L21 is read straight from S: it is only overwritten after the
    // barrier that follows the product loop.
    const int plast = ((n - 1) / PW) * PW;
    for (int p0 = plast; p0 >= 0; p0 -= PW) {
        double X[PW][PW];
        if (q == 0 && i >= p0) {
            double L[PW][PW];
#pragma unroll
            for (int r = 0; r < PW; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) L[r][c] = S[(p0 + r) * LDSM + p0 + c];
            trinv8(L, X);
        }
        double acc[PW];
#pragma unroll
        for (int c = 0; c < PW; ++c) acc[c] = 0.0;
        if (i >= p0 + PW && i < n) {
            for (int k = p0 + PW + q; k <= i; k += 2) {  // X22[i][k] (already inverted) times L21[k][:]
                const double x = S[i * LDSM + k];
                const double* lk = S + k * LDSM + p0;
#pragma unroll
                for (int c = 0; c < PW; ++c) acc[c] = fma(x, lk[c], acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < PW; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);  // q=0 + q=1 halves
        __syncthreads();  // every read of the old panel is done
        if (q == 0 && i >= p0 && i < n) {
            if (i >= p0 + PW) {
#pragma unroll
                for (int c = 0; c < PW; ++c) {
                    double t = 0.0;
#pragma unroll
                    for (int cc = c; cc < PW; ++cc) t = fma(acc[cc], X[cc][c], t);
                    S[i * LDSM + p0 + c] = -t;
                }
            } else {
#pragma unroll
                for (int r = 0; r < PW; ++r) {
                    if (i == p0 + r) {
#pragma unroll
                        for (int c = 0; c <= r; ++c) S[i * LDSM + p0 + c] = X[r][c];
                    }
                }
            }
        }
        __syncthreads();
    }

    {
        const int c = tid & (LEAF - 1), r0 = tid >> 7;
        if (c < n) {
#pragma unroll 8
            for (int r = r0; r < n; r += 2) {
                if (Dinv) Dinv[(int64_t)r * ldd + c] = (c <= r) ? S[r * LDSM + c] : 0.0;
                if (DinvT) DinvT[(int64_t)r * lddt + c] = (r <= c) ? S[c * LDSM + r] : 0.0;  // = Dinv[c][r]
            }
        }
    }
}

}  // namespace

int potrf_leaf(stream_t s, int n, double* A, int64_t lda, double* Dinv, int64_t ldd, double* DinvT, int64_t lddt,
               int* info, int64_t global_row0, int factor) {
    if (n < 0 || n > LEAF) return GPB_ERR_INVALID;
    if (n == 0) return GPB_OK;
    if (!A) return GPB_ERR_INVALID;
    constexpr size_t smem = sizeof(double) * (LEAF * LDSM);  // 129 KB: fits next to one resident GEMM CTA
    static PerDeviceOnce configured;
    const int dev = current_device();
    if (configured.needed(dev)) {
        if (cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
            return GPB_ERR_LAUNCH;
        configured.mark(dev);
    }
    potrf_leaf_kernel<<<1, LT, smem, to_stream(s)>>>(n, A, lda, Dinv, ldd, DinvT, lddt, info, global_row0, factor);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace gpb
