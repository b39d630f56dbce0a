// Single-CTA Cholesky + triangular inverse of one diagonal leaf block (n <= 128) held entirely in
// shared memory.  This is the latency-bound POTRF "panel" step of the blocked right-looking
// factorisation (reference: jnp.linalg.cholesky, gpjax/linalg/operations.py:55).  The explicit
// inverse it also emits turns every TRSM of the blocked algorithms into a DMMA GEMM.
//
// Failure semantics mirror JAX: a non-positive or NaN pivot NaN-fills the outputs; *info records
// the (1-based, global) index of the first failing pivot.
#include "common.cuh"

namespace gpb {

namespace {

constexpr int LEAF = 128;
constexpr int LDSM = LEAF + 1;  // odd stride: column walks are bank-conflict free
constexpr int LT = 512;         // threads: 4 per row

__global__ void __launch_bounds__(LT, 1) potrf_leaf_kernel(int n, double* __restrict__ A, int64_t lda,
                                                           double* __restrict__ Dinv, int64_t ldd,
                                                           double* __restrict__ DinvT, int64_t lddt, int* info,
                                                           int64_t global_row0, int factor) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                    // [LEAF][LDSM]
    double* col = sm + LEAF * LDSM;    // [LEAF]
    double* part = col + LEAF;         // [4][LEAF]
    __shared__ int fail;
    const int tid = threadIdx.x;
    const int i = tid & (LEAF - 1);  // row owned by this thread
    const int q = tid >> 7;          // which quarter of the columns it takes

    if (tid == 0) fail = 0;
    for (int idx = tid; idx < n * n; idx += LT) {
        int r = idx / n, c = idx % n;
        S[r * LDSM + c] = (c <= r) ? A[(int64_t)r * lda + c] : 0.0;
    }

    // ---- right-looking unblocked Cholesky --------------------------------------------------
    for (int j = 0; j < (factor ? n : 0); ++j) {
        __syncthreads();  // previous trailing update (and the load) complete
        double d = S[j * LDSM + j];
        if (!(d > 0.0)) {  // also catches NaN; uniform across the CTA
            if (tid == 0) fail = j + 1;
            break;
        }
        double p = sqrt(d);
        double inv = 1.0 / p;
        if (q == 0 && i > j && i < n) {
            double l = S[i * LDSM + j] * inv;
            S[i * LDSM + j] = l;
            col[i] = l;
        }
        __syncthreads();
        if (tid == 0) S[j * LDSM + j] = p;
        if (i > j && i < n) {
            double li = col[i];
            for (int c = j + 1 + q; c <= i; c += 4) S[i * LDSM + c] = fma(-li, col[c], S[i * LDSM + c]);
        }
    }
    __syncthreads();

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
    for (int idx = tid; factor && idx < n * n; idx += LT) {
        int r = idx / n, c = idx % n;
        if (c <= r) A[(int64_t)r * lda + c] = S[r * LDSM + c];
    }
    if (!Dinv && !DinvT) return;
    __syncthreads();

    // ---- in-place inverse of the lower-triangular factor (column sweep from the right) -------
    for (int j = n - 1; j >= 0; --j) {
        // v = L[j+1:, j] (old column), trailing block already holds inv(L[j+1:, j+1:])
        if (q == 0 && i > j && i < n) col[i] = S[i * LDSM + j];
        __syncthreads();
        double ajj = 1.0 / S[j * LDSM + j];
        if (i > j && i < n) {
            double acc = 0.0;
            for (int k = j + 1 + q; k <= i; k += 4) acc = fma(S[i * LDSM + k], col[k], acc);
            part[q * LEAF + i] = acc;
        }
        __syncthreads();
        if (q == 0) {
            if (i == j) S[j * LDSM + j] = ajj;
            else if (i > j && i < n)
                S[i * LDSM + j] = -ajj * (part[i] + part[LEAF + i] + part[2 * LEAF + i] + part[3 * LEAF + i]);
        }
        __syncthreads();
    }

    for (int idx = tid; idx < n * n; idx += LT) {
        int r = idx / n, c = idx % n;
        double v = (c <= r) ? S[r * LDSM + c] : 0.0;
        if (Dinv) Dinv[(int64_t)r * ldd + c] = v;
    }
    if (DinvT) {
        for (int idx = tid; idx < n * n; idx += LT) {
            int r = idx / n, c = idx % n;  // DinvT[r][c] = Dinv[c][r]
            DinvT[(int64_t)r * lddt + c] = (r <= c) ? S[c * LDSM + r] : 0.0;
        }
    }
}

}  // namespace

int potrf_leaf(stream_t s, int n, double* A, int64_t lda, double* Dinv, int64_t ldd, double* DinvT, int64_t lddt,
               int* info, int64_t global_row0, int factor) {
    if (n < 0 || n > LEAF) return GPB_ERR_INVALID;
    if (n == 0) return GPB_OK;
    if (!A) return GPB_ERR_INVALID;
    constexpr size_t smem = sizeof(double) * (LEAF * LDSM + LEAF + 4 * LEAF);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
            return GPB_ERR_LAUNCH;
        configured = true;
    }
    potrf_leaf_kernel<<<1, LT, smem, to_stream(s)>>>(n, A, lda, Dinv, ldd, DinvT, lddt, info, global_row0, factor);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace gpb
