// Small M x M / vector kernels of the collapsed-ELBO finish (gpjax/objectives.py:393-416 and the
// adjoints of SURVEY Appendix B).  All are bandwidth-trivial next to the streamed GEMMs.
#include "common.cuh"

namespace gpb {

namespace {

__global__ void set_identity_kernel(int64_t n, double* __restrict__ A, int64_t lda) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = blockIdx.y; r < n; r += gridDim.y)
        if (c < n) A[r * lda + c] = (r == c) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(1024) vec_sum_kernel(int64_t n, const double* __restrict__ x, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void scale_inplace_kernel(int64_t n, double* __restrict__ x, const double* __restrict__ f) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= f[0];
}

__global__ void axpy_kernel(int64_t n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fma(alpha, x[i], y[i]);
}

__global__ void aug_columns_kernel(int64_t rows, double* __restrict__ T, int64_t ld, int64_t M,
                                   const double* __restrict__ y, const double* __restrict__ c) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) {
        T[r * ld + M] = y[r] - (c ? c[0] : 0.0);
        T[r * ld + M + 1] = 1.0;
    }
}

__global__ void sgpr_prepare_kernel(int64_t M, const double* __restrict__ P, int64_t ldp,
                                    const double* __restrict__ obs_stddev, double* __restrict__ Bmat,
                                    double* __restrict__ psi, double* __restrict__ a1, double* __restrict__ sc) {
    const double sn = obs_stddev[0];
    const double s = sn * sn;
    const double is = 1.0 / s, isq = 1.0 / sqrt(s);
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = blockIdx.y; r < M; r += gridDim.y) {
        if (c < M) {
            double v = (c <= r) ? P[r * ldp + c] : P[c * ldp + r];  // lower stored
            Bmat[r * M + c] = v * is + ((r == c) ? 1.0 : 0.0);
        }
    }
    if (blockIdx.y == 0 && c < M) {
        psi[c] = P[M * ldp + c] * isq;
        a1[c] = P[(M + 1) * ldp + c] * isq;
    }
    if (blockIdx.y == 0 && blockIdx.x == 0) {
        // trace with a block reduction
        __shared__ double red[32];
        double t = 0.0;
        for (int64_t i = threadIdx.x; i < M; i += blockDim.x) t += P[i * ldp + i];
        t = block_sum(t, red);
        if (threadIdx.x == 0) {
            sc[0] = P[M * ldp + M];
            sc[1] = P[(M + 1) * ldp + M];
            sc[2] = P[(M + 1) * ldp + M + 1];
            sc[3] = t * is;
            sc[4] = s;
        }
    }
}

__global__ void sgpr_value_kernel(const double* sc, const double* hl, const double* wtw, const double* variance,
                                  const int* info, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double dd = sc[0], n = sc[2], trphi = sc[3], s = sc[4];
        double two_log_prob = -n * log(6.283185307179586 * s) - 2.0 * hl[0] - (dd - wtw[0]) / s;
        double two_trace = n * variance[0] / s - trphi;
        double v = 0.5 * (two_log_prob - two_trace);
        if (info && (info[0] != 0 || info[1] != 0)) v = nan("");
        out[0] = v;
    }
}

__global__ void __launch_bounds__(256) sgpr_adjoints_kernel(int64_t M, const double* __restrict__ Binv,
                                                            const double* __restrict__ Bmat,
                                                            const double* __restrict__ v, const double* __restrict__ sc,
                                                            double* __restrict__ G1, double* __restrict__ G2,
                                                            double* __restrict__ u, double* __restrict__ rowsum) {
    __shared__ double red[32];
    const double s = sc[4];
    const int64_t r = blockIdx.x;
    const double vr = v[r];
    double acc = 0.0;
    for (int64_t c = threadIdx.x; c < M; c += blockDim.x) {
        double eye = (r == c) ? 1.0 : 0.0;
        double dphi = 0.5 * (eye - Binv[r * M + c] - vr * v[c] / s);
        double phi = Bmat[r * M + c] - eye;
        G1[r * M + c] = (2.0 / s) * dphi;
        G2[r * M + c] = dphi - 0.5 * phi;
        acc = fma(dphi, phi, acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        rowsum[r] = acc;
        u[r] = vr / (s * sqrt(s));
    }
}

__global__ void sgpr_scalar_grads_kernel(const double* sc, const double* dots, const double* variance,
                                         const double* obs_stddev, double* g_var, double* g_obs, double* g_mean) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double dd = sc[0], sd = sc[1], n = sc[2], s = sc[4];
        const double psiv = dots[0], va1 = dots[1], dpp = dots[2];
        const double var = variance[0], sn = obs_stddev[0];
        double g_s = -n / (2.0 * s) + (dd - psiv) / (2.0 * s * s) + n * var / (2.0 * s * s) -
                     (2.0 * dpp + psiv / s) / (2.0 * s);
        if (g_var) g_var[0] += -n / (2.0 * s);
        if (g_obs) g_obs[0] = 2.0 * sn * g_s;
        if (g_mean) g_mean[0] = -va1 / s + sd / s;
    }
}

}  // namespace

int set_identity(stream_t s, int64_t n, double* A, int64_t lda) {
    if (n <= 0) return GPB_OK;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(n < 4096 ? n : 4096));
    set_identity_kernel<<<grid, 256, 0, to_stream(s)>>>(n, A, lda);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int vec_sum(stream_t s, int64_t n, const double* x, double* out) {
    vec_sum_kernel<<<1, 1024, 0, to_stream(s)>>>(n, x, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int scale_inplace(stream_t s, int64_t n, double* x, const double* f) {
    if (n <= 0 || !f || !x) return GPB_OK;
    scale_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, to_stream(s)>>>(n, x, f);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int axpy(stream_t s, int64_t n, double alpha, const double* x, double* y) {
    if (n <= 0) return GPB_OK;
    axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, to_stream(s)>>>(n, alpha, x, y);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int sgpr_aug_columns(stream_t s, int64_t rows, double* T, int64_t ld, int64_t M, const double* y, const double* c) {
    if (rows <= 0) return GPB_OK;
    aug_columns_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, to_stream(s)>>>(rows, T, ld, M, y, c);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int sgpr_prepare(stream_t s, int64_t M, const double* Paug, int64_t ldp, const double* obs_stddev, double* Bmat,
                 double* psi, double* a1, double* sc) {
    if (M <= 0) return GPB_ERR_INVALID;
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)(M < 4096 ? M : 4096));
    sgpr_prepare_kernel<<<grid, 256, 0, to_stream(s)>>>(M, Paug, ldp, obs_stddev, Bmat, psi, a1, sc);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int sgpr_value(stream_t s, const double* sc, const double* half_logdetB, const double* wtw, const double* variance,
               const int* info, double* out) {
    sgpr_value_kernel<<<1, 32, 0, to_stream(s)>>>(sc, half_logdetB, wtw, variance, info, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int sgpr_adjoints(stream_t s, int64_t M, const double* Binv, const double* Bmat, const double* v, const double* sc,
                  double* G1, double* G2, double* u, double* rowsum) {
    if (M <= 0) return GPB_ERR_INVALID;
    sgpr_adjoints_kernel<<<(unsigned)M, 256, 0, to_stream(s)>>>(M, Binv, Bmat, v, sc, G1, G2, u, rowsum);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int sgpr_scalar_grads(stream_t s, const double* sc, const double* dots, const double* variance,
                      const double* obs_stddev, double* g_var, double* g_obs, double* g_mean) {
    sgpr_scalar_grads_kernel<<<1, 32, 0, to_stream(s)>>>(sc, dots, variance, obs_stddev, g_var, g_obs, g_mean);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace gpb
