// Small M x M / vector kernels of the SVGP (uncollapsed) ELBO finish: gpjax/objectives.py:241-315,
// variational_families.py:169-285, distributions.py:188-228, integrators.py:151-158 reduced to the same
// row-additive statistics as the collapsed bound.  Bandwidth-trivial.
#include "common.cuh"

namespace gpb {

namespace {

__global__ void __launch_bounds__(1024) sum_log_abs_diag_kernel(int64_t n, const double* __restrict__ A, int64_t lda,
                                                                double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += log(fabs(A[i * (lda + 1)]));
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void svgp_unpack_kernel(int64_t M, const double* __restrict__ P, int64_t ldp,
                                   const double* __restrict__ obs_stddev, double num_datapoints,
                                   double* __restrict__ Phi, double* __restrict__ psi, double* __restrict__ a1,
                                   double* __restrict__ sc) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = blockIdx.y; r < M; r += gridDim.y)
        if (c < M) Phi[r * M + c] = (c <= r) ? P[r * ldp + c] : P[c * ldp + r];
    if (blockIdx.y == 0 && c < M) {
        psi[c] = P[M * ldp + c];
        a1[c] = P[(M + 1) * ldp + c];
    }
    if (blockIdx.y == 0 && blockIdx.x == 0) {
        __shared__ double red[32];
        double t = 0.0;
        for (int64_t i = threadIdx.x; i < M; i += blockDim.x) t += P[i * ldp + i];
        t = block_sum(t, red);
        if (threadIdx.x == 0) {
            const double sn = obs_stddev[0], s = sn * sn, B = P[(M + 1) * ldp + M + 1];
            sc[0] = P[M * ldp + M];
            sc[1] = P[(M + 1) * ldp + M];
            sc[2] = B;
            sc[3] = t;
            sc[4] = s;
            sc[5] = (num_datapoints / B) / s;
        }
    }
}

__global__ void svgp_value_kernel(int64_t M, const double* sc, const double* dots, const double* variance, double jitter,
                                  const int* info, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double dd = sc[0], B = sc[2], trphi = sc[3], s = sc[4], coef = sc[5];
        const double Q = dd - 2.0 * dots[0] + dots[3] + B * (variance[0] + jitter) - trphi;
        const double ell = -0.5 * (B * log(6.283185307179586 * s) + Q / s);
        const double kl = 0.5 * (dots[1] - (double)M - 2.0 * dots[5] + 2.0 * dots[4] + dots[2]);
        double v = coef * s * ell - kl;  // coef * s = N / B
        if (info && info[0] != 0) v = nan("");
        out[0] = v;
    }
}

__global__ void svgp_adjoints_kernel(int64_t M, const double* __restrict__ Phi, const double* __restrict__ Tt,
                                     const double* __restrict__ PT, const double* __restrict__ u,
                                     const double* __restrict__ psi, const double* __restrict__ sc,
                                     double* __restrict__ G1, double* __restrict__ E) {
    const double coef = sc[5];
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = blockIdx.y; r < M; r += gridDim.y) {
        if (c < M) {
            const double eye = (r == c) ? 1.0 : 0.0;
            const double tt = Tt[r * M + c];
            G1[r * M + c] = coef * (eye - tt);
            E[r * M + c] = -0.5 * coef * (u[r] * psi[c] + u[c] * psi[r]) + 0.5 * coef * (PT[r * M + c] + PT[c * M + r]) -
                           0.5 * coef * Phi[r * M + c] + 0.5 * tt - 0.5 * eye;
        }
    }
}

__global__ void svgp_vectors_kernel(int64_t M, const double* psi, const double* Phiu, const double* u, const double* sc,
                                    double* tvec, double* uvec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
        const double coef = sc[5];
        tvec[i] = coef * (psi[i] - Phiu[i]) - u[i];
        uvec[i] = coef * u[i];
    }
}

__global__ void svgp_h_kernel(int64_t n, const double* __restrict__ PhiV, const double* __restrict__ V,
                              const double* __restrict__ sc, double* __restrict__ H) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) H[i] = fma(sc[5], PhiV[i], V[i]);
}

__global__ void svgp_gw_diag_kernel(int64_t M, const double* W, int64_t ldw, double* gW, int64_t ldg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) gW[i * ldg + i] += 1.0 / W[i * ldw + i];
}

__global__ void svgp_scalar_grads_kernel(const double* sc, const double* dots, const double* dots2,
                                         const double* variance, const double* obs_stddev, double jitter, double* g_var,
                                         double* g_obs, double* g_mean) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double dd = sc[0], sd = sc[1], B = sc[2], trphi = sc[3], s = sc[4], coef = sc[5];
        const double kappa = coef * s;
        const double Q = dd - 2.0 * dots[0] + dots[3] + B * (variance[0] + jitter) - trphi;
        if (g_var) g_var[0] += -kappa * B / (2.0 * s);
        if (g_obs) g_obs[0] = 2.0 * obs_stddev[0] * kappa * (-B / (2.0 * s) + Q / (2.0 * s * s));
        if (g_mean) g_mean[0] = coef * (sd - dots2[0]) - dots2[1];
    }
}

}  // namespace

int sum_log_abs_diag(stream_t s, int64_t n, const double* A, int64_t lda, double* out) {
    sum_log_abs_diag_kernel<<<1, 1024, 0, to_stream(s)>>>(n, A, lda, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int svgp_unpack(stream_t s, int64_t M, const double* Paug, int64_t ldp, const double* obs_stddev, double num_datapoints,
                double* Phi, double* psi, double* a1, double* sc) {
    if (M <= 0) return GPB_ERR_INVALID;
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)(M < 4096 ? M : 4096));
    svgp_unpack_kernel<<<grid, 256, 0, to_stream(s)>>>(M, Paug, ldp, obs_stddev, num_datapoints, Phi, psi, a1, sc);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int svgp_value(stream_t s, int64_t M, const double* sc, const double* dots, const double* variance, double jitter,
               const int* info, double* out) {
    svgp_value_kernel<<<1, 32, 0, to_stream(s)>>>(M, sc, dots, variance, jitter, info, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int svgp_adjoints(stream_t s, int64_t M, const double* Phi, const double* Ttil, const double* PT, const double* u,
                  const double* psi, const double* sc, double* G1, double* E) {
    if (M <= 0) return GPB_ERR_INVALID;
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)(M < 4096 ? M : 4096));
    svgp_adjoints_kernel<<<grid, 256, 0, to_stream(s)>>>(M, Phi, Ttil, PT, u, psi, sc, G1, E);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int svgp_vectors(stream_t s, int64_t M, const double* psi, const double* Phiu, const double* u, const double* sc,
                 double* tvec, double* uvec) {
    svgp_vectors_kernel<<<(unsigned)((M + 255) / 256), 256, 0, to_stream(s)>>>(M, psi, Phiu, u, sc, tvec, uvec);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int svgp_h(stream_t s, int64_t M, const double* PhiV, const double* V, const double* sc, double* H) {
    const int64_t n = M * M;
    svgp_h_kernel<<<(unsigned)((n + 255) / 256), 256, 0, to_stream(s)>>>(n, PhiV, V, sc, H);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int svgp_gw_diag(stream_t s, int64_t M, const double* W, int64_t ldw, double* gW, int64_t ldg) {
    svgp_gw_diag_kernel<<<(unsigned)((M + 255) / 256), 256, 0, to_stream(s)>>>(M, W, ldw, gW, ldg);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int svgp_scalar_grads(stream_t s, const double* sc, const double* dots, const double* dots2, const double* variance,
                      const double* obs_stddev, double jitter, double* g_var, double* g_obs, double* g_mean) {
    svgp_scalar_grads_kernel<<<1, 32, 0, to_stream(s)>>>(sc, dots, dots2, variance, obs_stddev, jitter, g_var, g_obs,
                                                         g_mean);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace gpb
