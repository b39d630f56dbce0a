// Bandwidth-bound helpers: GEMV (triangular-solve building block), reductions, copies, scalar assembly.
// Reference call sites these stand in for: jsp.linalg.solve_triangular with a vector RHS
// (gpjax/linalg/operations.py:105-107), sum(log(diag)) (operations.py:142-144), diff.T @ solve(...)
// (gpjax/distributions.py:132-134).
#include "common.cuh"

namespace gpb {

namespace {

// y[m] = beta*y + alpha * A[m x n] x : one warp per row, lanes stride the row (coalesced).
__global__ void __launch_bounds__(256) gemv_n_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda,
                                                     const double* __restrict__ x, double* __restrict__ y,
                                                     double alpha, double beta) {
    int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= m) return;
    const double* a = A + row * lda;
    double s = 0.0;
    for (int64_t j = lane; j < n; j += 32) s = fma(a[j], x[j], s);
    s = warp_sum(s);
    if (lane == 0) y[row] = (beta == 0.0 ? 0.0 : beta * y[row]) + alpha * s;
}

// y[n] = beta*y + alpha * A[m x n]^T x : a CTA owns 64 columns; its 4 thread groups take every 4th row
// (coalesced 512-byte row segments, 4-way unrolled for memory-level parallelism) and combine in smem.
__global__ void __launch_bounds__(256) gemv_t_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda,
                                                     const double* __restrict__ x, double* __restrict__ y,
                                                     double alpha, double beta) {
    __shared__ double red[4][64];
    const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int64_t col = (int64_t)blockIdx.x * 64 + cg;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (col < n) {
        const double* a = A + col;
        int64_t r = rg;
        for (; r + 12 < m; r += 16) {
            s0 = fma(a[r * lda], x[r], s0);
            s1 = fma(a[(r + 4) * lda], x[r + 4], s1);
            s2 = fma(a[(r + 8) * lda], x[r + 8], s2);
            s3 = fma(a[(r + 12) * lda], x[r + 12], s3);
        }
        for (; r < m; r += 4) s0 = fma(a[r * lda], x[r], s0);
    }
    red[rg][cg] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (rg == 0 && col < n) {
        double s = (red[0][cg] + red[1][cg]) + (red[2][cg] + red[3][cg]);
        y[col] = (beta == 0.0 ? 0.0 : beta * y[col]) + alpha * s;
    }
}

__global__ void __launch_bounds__(1024) sum_log_diag_kernel(int64_t n, const double* __restrict__ A, int64_t lda,
                                                            double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += log(A[i * (lda + 1)]);
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(1024) dot_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                   double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fma(x[i], y[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

// Long dot products (the M x M Frobenius terms of the SVGP finish: 16.8 M entries took 1 ms on one CTA): one portable cluster of
// 8 CTAs, each sums a strided eighth, rank 0 adds the eight partial sums through distributed shared memory in rank order --
// deterministic, no scratch buffer, no atomics.
constexpr int DOT_CLUSTER = 8;
__global__ void __cluster_dims__(DOT_CLUSTER, 1, 1) __launch_bounds__(1024)
    dot_cluster_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y, double* out) {
    __shared__ double red[32];
    __shared__ double part;
    unsigned rank, peer_addr;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    double s = 0.0;
    for (int64_t i = (int64_t)rank * 1024 + threadIdx.x; i < n; i += (int64_t)DOT_CLUSTER * 1024) s = fma(x[i], y[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) part = s;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank == 0 && threadIdx.x == 0) {
        double t = 0.0;
        const unsigned local = (unsigned)__cvta_generic_to_shared(&part);
        for (unsigned r = 0; r < DOT_CLUSTER; ++r) {
            double v;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_addr) : "r"(local), "r"(r));
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(peer_addr) : "memory");
            t += v;
        }
        out[0] = t;
    }
    // no CTA may exit (and release its shared memory) before rank 0 has read every partial sum
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void sub_scalar_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ c,
                                  double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double cv = c ? c[0] : 0.0;
    if (i < n) out[i] = a[i] - cv;
}

__global__ void transpose_kernel(int64_t rows, int64_t cols, const double* __restrict__ src, int64_t lds,
                                 double* __restrict__ dst, int64_t ldd) {
    __shared__ double t[32][33];
    int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t r = (int64_t)blockIdx.y * 32 + k;
        if (r < rows && c < cols) t[k][threadIdx.x] = src[r * lds + c];
    }
    __syncthreads();
    int64_t r2 = (int64_t)blockIdx.y * 32 + threadIdx.x;  // original row -> new column
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t c2 = (int64_t)blockIdx.x * 32 + k;  // original col -> new row
        if (r2 < rows && c2 < cols) dst[c2 * ldd + r2] = t[threadIdx.x][k];
    }
}

__global__ void fill2d_kernel(int64_t rows, int64_t cols, double* __restrict__ p, int64_t ld, double v) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = blockIdx.y;
    for (; r < rows; r += gridDim.y)
        if (c < cols) p[r * ld + c] = v;
}

__global__ void zero_triangle_kernel(int64_t n, double* __restrict__ A, int64_t lda, int uplo) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
        if (c < n) {
            if ((uplo == 2 && c > r) || (uplo == 1 && c < r)) A[r * lda + c] = 0.0;
        }
    }
}

// tile-transposing mirror so both the read and the write side are coalesced
__global__ void symmetrize_kernel(int64_t n, double* __restrict__ A, int64_t lda, int from_lower) {
    __shared__ double t[32][33];
    int64_t bi = blockIdx.y, bj = blockIdx.x;  // tile (bi, bj) with bi >= bj is a source tile (lower)
    if (bj > bi) return;
    // source tile coordinates in the triangle we read from
    int64_t sr0 = (from_lower ? bi : bj) * 32, sc0 = (from_lower ? bj : bi) * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t r = sr0 + k, c = sc0 + threadIdx.x;
        if (r < n && c < n) t[k][threadIdx.x] = A[r * lda + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t r = sc0 + k, c = sr0 + threadIdx.x;  // destination = transposed position
        if (r < n && c < n) {
            bool strictly_other = from_lower ? (c > r) : (c < r);
            if (strictly_other) A[r * lda + c] = t[threadIdx.x][k];
        }
    }
}

// lower triangle <- (A + A^T) / 2 (what jnp.linalg.cholesky does to its input before factorising: symmetrize_input=True);
// tile (bi, bj), bi >= bj: the mirrored upper tile is read row-wise and transposed through shared memory
__global__ void sym_average_lower_kernel(int64_t n, double* __restrict__ A, int64_t lda) {
    __shared__ double t[32][33];
    const int64_t bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {  // upper tile (bj, bi)
        const int64_t r = bj * 32 + k, c = bi * 32 + threadIdx.x;
        if (r < n && c < n) t[k][threadIdx.x] = A[r * lda + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {  // lower tile (bi, bj)
        const int64_t r = bi * 32 + k, c = bj * 32 + threadIdx.x;
        if (r < n && c < r) A[r * lda + c] = 0.5 * (A[r * lda + c] + t[threadIdx.x][k]);
    }
}

__global__ void mll_value_kernel(int64_t n, const double* half_logdet, const double* quad, const int* info,
                                 double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double v = -0.5 * ((double)n * 1.8378770664093453 + 2.0 * half_logdet[0] + quad[0]);
        if (info && info[0] != 0) v = nan("");
        out[0] = v;
    }
}

}  // namespace

int gemv(stream_t s, int64_t m, int64_t n, const double* A, int64_t lda, int trans, const double* x, double* y,
         double alpha, double beta) {
    if (m < 0 || n < 0) return GPB_ERR_INVALID;
    cudaStream_t st = to_stream(s);
    if (trans == 0) {
        if (m == 0) return GPB_OK;
        gemv_n_kernel<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(m, n, A, lda, x, y, alpha, beta);
    } else {
        if (n == 0) return GPB_OK;
        gemv_t_kernel<<<(unsigned)((n + 63) / 64), 256, 0, st>>>(m, n, A, lda, x, y, alpha, beta);
    }
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int sum_log_diag(stream_t s, int64_t n, const double* A, int64_t lda, double* out) {
    sum_log_diag_kernel<<<1, 1024, 0, to_stream(s)>>>(n, A, lda, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int dot(stream_t s, int64_t n, const double* x, const double* y, double* out) {
    if (n >= (1 << 17)) dot_cluster_kernel<<<DOT_CLUSTER, 1024, 0, to_stream(s)>>>(n, x, y, out);
    else dot_kernel<<<1, 1024, 0, to_stream(s)>>>(n, x, y, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int sub_scalar(stream_t s, int64_t n, const double* a, const double* c, double* out) {
    if (n <= 0) return GPB_OK;
    sub_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, to_stream(s)>>>(n, a, c, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int copy2d(stream_t s, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst, int64_t ldd) {
    if (rows <= 0 || cols <= 0) return GPB_OK;
    cudaError_t e = cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), cols * sizeof(double),
                                      rows, cudaMemcpyDeviceToDevice, to_stream(s));
    return e == cudaSuccess ? GPB_OK : GPB_ERR_LAUNCH;
}

int transpose2d(stream_t s, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst, int64_t ldd) {
    if (rows <= 0 || cols <= 0) return GPB_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, to_stream(s)>>>(rows, cols, src, lds, dst, ldd);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int fill2d(stream_t s, int64_t rows, int64_t cols, double* p, int64_t ld, double v) {
    if (rows <= 0 || cols <= 0) return GPB_OK;
    dim3 grid((unsigned)((cols + 255) / 256), (unsigned)(rows < 4096 ? rows : 4096));
    fill2d_kernel<<<grid, 256, 0, to_stream(s)>>>(rows, cols, p, ld, v);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int zero_triangle(stream_t s, int64_t n, double* A, int64_t lda, int uplo) {
    if (n <= 0) return GPB_OK;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(n < 4096 ? n : 4096));
    zero_triangle_kernel<<<grid, 256, 0, to_stream(s)>>>(n, A, lda, uplo);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int symmetrize(stream_t s, int64_t n, double* A, int64_t lda, int from_lower) {
    if (n <= 0) return GPB_OK;
    unsigned t = (unsigned)((n + 31) / 32);
    symmetrize_kernel<<<dim3(t, t), dim3(32, 8), 0, to_stream(s)>>>(n, A, lda, from_lower);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int symmetrize_average_lower(stream_t s, int64_t n, double* A, int64_t lda) {
    if (n <= 0) return GPB_OK;
    unsigned t = (unsigned)((n + 31) / 32);
    sym_average_lower_kernel<<<dim3(t, t), dim3(32, 8), 0, to_stream(s)>>>(n, A, lda);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int mll_value(stream_t s, int64_t n, const double* half_logdet, const double* quad, const int* info, double* out) {
    mll_value_kernel<<<1, 32, 0, to_stream(s)>>>(n, half_logdet, quad, info, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace gpb
