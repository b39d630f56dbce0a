#include <cstdlib>
// HOST MODEL of gpjax_b200/csrc/primitives.h -- TEST INFRASTRUCTURE ONLY.
//
// Plain C++ loops with the same argument semantics as the CUDA launchers, operating on host
// pointers and ignoring the stream.  Linked with the *product's* algorithms.cpp/abi.cpp into
// tests/hostsim/libgpjax_b200_hostsim.so so the CPU test-suite can check the blocked-algorithm
// orchestration (block indexing, masks, workspace carving, K-range hints) against the oracle
// without a GPU.  It is never shipped, never loaded by gpjax_b200, and is not a fallback.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#include "../../gpjax_b200/csrc/primitives.h"

namespace gpb {

static inline double elemA(const GemmDesc& d, const double* A, int64_t m, int64_t k) {
    return d.a_layout == LAYOUT_K ? A[m * d.lda + k] : A[k * d.lda + m];
}
static inline double elemB(const GemmDesc& d, const double* B, int64_t n, int64_t k) {
    return d.b_layout == LAYOUT_K ? B[n * d.ldb + k] : B[k * d.ldb + n];
}
static inline bool keep(const GemmDesc& d, int64_t r, int64_t c) {
    switch (d.mask) {
        case MASK_LOWER: return r >= c;
        case MASK_UPPER: return r <= c;
        case MASK_BLOCK_STRICT_UPPER: return (r / d.mask_nb) < (c / d.mask_nb);
        case MASK_BLOCK_STRICT_LOWER: return (r / d.mask_nb) > (c / d.mask_nb);
        default: return true;
    }
}

int gemm(stream_t, const GemmDesc& d) {
    if (d.M < 0 || d.N < 0 || d.K < 0) return GPB_ERR_INVALID;
    // Same tile geometry as the CUDA kernel so the K-range hints are exercised identically:
    // a hint that would skip a physically non-zero element shows up as a wrong answer here too.
    const int64_t BM = 128, BN = 64, BK = 16;
    for (int b = 0; b < d.batch; ++b) {
        const double* A = d.A + b * d.strideA;
        const double* B = d.B + b * d.strideB;
        double* C = d.C + b * d.strideC;
        for (int64_t m0 = 0; m0 < d.M; m0 += BM)
            for (int64_t n0 = 0; n0 < d.N; n0 += BN) {
                int64_t mend = std::min(m0 + BM, d.M), nend = std::min(n0 + BN, d.N);
                int64_t kbeg = 0, kend = d.K;
                if (d.krange == KR_B_LOWER) kend = std::min(d.K, nend + d.kr_off);
                else if (d.krange == KR_B_UPPER) kbeg = std::max<int64_t>(0, n0 + d.kr_off);
                else if (d.krange == KR_A_LOWER) kend = std::min(d.K, mend + d.kr_off);
                else if (d.krange == KR_A_UPPER) kbeg = std::max<int64_t>(0, m0 + d.kr_off);
                kbeg = (kbeg / BK) * BK;
                if (kend < kbeg) kend = kbeg;
                for (int64_t m = m0; m < mend; ++m)
                    for (int64_t n = n0; n < nend; ++n) {
                        if (!keep(d, d.mask_row0 + m, d.mask_col0 + n)) continue;
                        double acc = 0.0;
                        for (int64_t k = kbeg; k < kend; ++k) acc += elemA(d, A, m, k) * elemB(d, B, n, k);
                        double v = d.alpha * acc;
                        if (d.beta != 0.0) v += d.beta * C[m * d.ldc + n];
                        C[m * d.ldc + n] = v;
                    }
            }
    }
    return GPB_OK;
}

// One kernel evaluation with all first derivatives, written pair-by-pair (an independent scalar model of the
// tiled device code).  `scal` = {variance, shape}: shape = alpha (RationalQuadratic), power (PoweredExponential),
// period (Periodic).
struct PairOut {
    double k = 0.0;        // value
    double dshape = 0.0;   // d k / d shape
    double dx[64];         // d k / d x_d   (= - d k / d z_d)
    double dl[64];         // d k / d l_d
};
static void pair_eval(int kind, const double* x, const double* z, int D, const double* ell, int iso,
                      const double* scal, PairOut& o) {
    const double var = scal[0];
    const double shp = kind_has_shape(kind) ? scal[1] : 0.0;
    o.dshape = 0.0;
    if (kind == KIND_PERIODIC) {
        const double pi = 3.141592653589793;
        double r2 = 0.0, s[64], c[64], u[64];
        for (int d = 0; d < D; ++d) {
            u[d] = pi * (x[d] - z[d]) / shp;
            s[d] = std::sin(u[d]) / ell[iso ? 0 : d];
            c[d] = std::cos(u[d]);
            r2 += s[d] * s[d];
        }
        o.k = var * std::exp(-0.5 * r2);
        for (int d = 0; d < D; ++d) {
            double l = ell[iso ? 0 : d];
            o.dx[d] = -o.k * s[d] * c[d] / l * pi / shp;
            o.dl[d] = o.k * s[d] * s[d] / l;
            o.dshape += o.k * s[d] * c[d] / l * u[d] / shp;
        }
        return;
    }
    double r2 = 0.0, a[64];
    for (int d = 0; d < D; ++d) {
        double l = ell[iso ? 0 : d];
        a[d] = x[d] / l - z[d] / l;
        r2 += a[d] * a[d];
    }
    double k, dk = 0.0;  // dk = d k / d r2
    const double tau = std::sqrt(std::fmax(r2, 1e-36));
    const bool live = r2 > 1e-36;
    switch (kind) {
        case KIND_RBF: k = var * std::exp(-0.5 * r2); dk = -0.5 * k; break;
        case KIND_MATERN12: k = var * std::exp(-tau); dk = live ? -0.5 * k / tau : 0.0; break;
        case KIND_MATERN32:
            k = var * (1.0 + std::sqrt(3.0) * tau) * std::exp(-std::sqrt(3.0) * tau);
            dk = live ? -1.5 * var * std::exp(-std::sqrt(3.0) * tau) : 0.0;
            break;
        case KIND_MATERN52:
            k = var * (1.0 + std::sqrt(5.0) * tau + 5.0 / 3.0 * tau * tau) * std::exp(-std::sqrt(5.0) * tau);
            dk = live ? -(5.0 / 6.0) * var * (1.0 + std::sqrt(5.0) * tau) * std::exp(-std::sqrt(5.0) * tau) : 0.0;
            break;
        case KIND_RATQUAD: {
            double base = 1.0 + 0.5 * r2 / shp;
            k = var * std::pow(base, -shp);
            dk = -0.5 * var * std::pow(base, -shp - 1.0);
            o.dshape = k * (-std::log(base) + (0.5 * r2 / shp) / base);
            break;
        }
        case KIND_POWEXP: {
            double tp = std::pow(tau, shp);
            k = var * std::exp(-tp);
            dk = live ? -k * shp * std::pow(tau, shp - 1.0) / (2.0 * tau) : 0.0;
            o.dshape = -k * tp * std::log(tau);
            break;
        }
        default: k = (r2 == 0.0) ? var : 0.0; dk = 0.0; break;  // KIND_WHITE
    }
    o.k = k;
    for (int d = 0; d < D; ++d) {
        double l = ell[iso ? 0 : d];
        o.dx[d] = dk * 2.0 * a[d] / l;
        o.dl[d] = dk * (-2.0) * a[d] * a[d] / l;
    }
}

int max_input_dim() { return 64; }

int gram(stream_t, const GramDesc& d) {
    if (d.N < 0 || d.M < 0 || d.D <= 0) return GPB_ERR_INVALID;
    if (d.D > 64) return GPB_ERR_UNSUPPORTED;
    if (!kind_valid(d.kind)) return GPB_ERR_INVALID;
    double dadd = d.diag_add + (d.diag_add_sq ? d.diag_add_sq[0] * d.diag_add_sq[0] : 0.0);
    const int64_t TR = 64, TC = 128;
    for (int64_t i = 0; i < d.N; ++i)
        for (int64_t j = 0; j < d.M; ++j) {
            if (d.lower_only) {  // tile-level skip, as on the device
                int64_t r0 = (i / TR) * TR, c0 = (j / TC) * TC;
                if (d.row0 + std::min(r0 + TR, d.N) - 1 < d.col0 + c0) continue;
            }
            PairOut po;
            pair_eval(d.kind, d.X + i * d.ldx, d.Z + j * d.ldz, d.D, d.ell, d.ell_is_scalar, d.variance, po);
            double k = po.k;
            if (d.row0 + i == d.col0 + j) k += dadd;
            d.K[i * d.ldk + j] = k;
        }
    return GPB_OK;
}

int potrf_leaf(stream_t, int n, double* A, int64_t lda, double* Dinv, int64_t ldd, double* DinvT, int64_t lddt,
               int* info, int64_t global_row0, int factor) {
    if (n < 0 || n > 128) return GPB_ERR_INVALID;
    std::vector<double> S((size_t)n * n, 0.0);
    for (int r = 0; r < n; ++r)
        for (int c = 0; c <= r; ++c) S[r * n + c] = A[r * lda + c];
    int fail = 0;
    if (factor) {
        for (int j = 0; j < n && !fail; ++j) {
            double dj = S[j * n + j];
            if (!(dj > 0.0)) { fail = j + 1; break; }
            double p = std::sqrt(dj), inv = 1.0 / p;
            S[j * n + j] = p;
            for (int i = j + 1; i < n; ++i) S[i * n + j] *= inv;
            for (int i = j + 1; i < n; ++i)
                for (int c = j + 1; c <= i; ++c) S[i * n + c] -= S[i * n + j] * S[c * n + j];
        }
    }
    if (fail) {
        double q = std::numeric_limits<double>::quiet_NaN();
        if (info && *info == 0) *info = (int)(global_row0 + fail);
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c) {
                if (c <= r) A[r * lda + c] = q;
                if (Dinv) Dinv[r * ldd + c] = q;
                if (DinvT) DinvT[r * lddt + c] = q;
            }
        return GPB_OK;
    }
    if (factor)
        for (int r = 0; r < n; ++r)
            for (int c = 0; c <= r; ++c) A[r * lda + c] = S[r * n + c];
    // inverse by forward substitution on unit vectors
    std::vector<double> X((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j)
        for (int i = j; i < n; ++i) {
            double acc = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) acc -= S[i * n + k] * X[k * n + j];
            X[i * n + j] = acc / S[i * n + i];
        }
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            if (Dinv) Dinv[r * ldd + c] = X[r * n + c];
            if (DinvT) DinvT[r * lddt + c] = X[c * n + r];
        }
    return GPB_OK;
}

int gemv(stream_t, int64_t m, int64_t n, const double* A, int64_t lda, int trans, const double* x, double* y,
         double alpha, double beta) {
    if (!trans) {
        for (int64_t i = 0; i < m; ++i) {
            double s = 0.0;
            for (int64_t j = 0; j < n; ++j) s += A[i * lda + j] * x[j];
            y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * s;
        }
    } else {
        for (int64_t j = 0; j < n; ++j) {
            double s = 0.0;
            for (int64_t i = 0; i < m; ++i) s += A[i * lda + j] * x[i];
            y[j] = (beta == 0.0 ? 0.0 : beta * y[j]) + alpha * s;
        }
    }
    return GPB_OK;
}

int sum_log_diag(stream_t, int64_t n, const double* A, int64_t lda, double* out) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += std::log(A[i * (lda + 1)]);
    out[0] = s;
    return GPB_OK;
}
int dot(stream_t, int64_t n, const double* x, const double* y, double* out) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    out[0] = s;
    return GPB_OK;
}
int sub_scalar(stream_t, int64_t n, const double* a, const double* c, double* out) {
    double cv = c ? c[0] : 0.0;
    for (int64_t i = 0; i < n; ++i) out[i] = a[i] - cv;
    return GPB_OK;
}
int copy2d(stream_t, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst, int64_t ldd) {
    for (int64_t i = 0; i < rows; ++i) std::memmove(dst + i * ldd, src + i * lds, (size_t)cols * sizeof(double));
    return GPB_OK;
}
int transpose2d(stream_t, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst, int64_t ldd) {
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) dst[j * ldd + i] = src[i * lds + j];
    return GPB_OK;
}
int fill2d(stream_t, int64_t rows, int64_t cols, double* p, int64_t ld, double v) {
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) p[i * ld + j] = v;
    return GPB_OK;
}
int zero_triangle(stream_t, int64_t n, double* A, int64_t lda, int uplo) {
    for (int64_t r = 0; r < n; ++r)
        for (int64_t c = 0; c < n; ++c)
            if ((uplo == 2 && c > r) || (uplo == 1 && c < r)) A[r * lda + c] = 0.0;
    return GPB_OK;
}
int symmetrize(stream_t, int64_t n, double* A, int64_t lda, int from_lower) {
    for (int64_t r = 0; r < n; ++r)
        for (int64_t c = 0; c < r; ++c) {
            if (from_lower) A[c * lda + r] = A[r * lda + c];
            else A[r * lda + c] = A[c * lda + r];
        }
    return GPB_OK;
}
int symmetrize_average_lower(stream_t, int64_t n, double* A, int64_t lda) {
    for (int64_t r = 0; r < n; ++r)
        for (int64_t c = 0; c < r; ++c) A[r * lda + c] = 0.5 * (A[r * lda + c] + A[c * lda + r]);
    return GPB_OK;
}
int mll_value(stream_t, int64_t n, const double* half_logdet, const double* quad, const int* info, double* out) {
    double v = -0.5 * ((double)n * std::log(2.0 * M_PI) + 2.0 * half_logdet[0] + quad[0]);
    if (info && info[0] != 0) v = std::numeric_limits<double>::quiet_NaN();
    out[0] = v;
    return GPB_OK;
}

int64_t mll_bwd_partials_count(int64_t N, int D, int64_t) {
    int DC = D <= 2 ? 2 : (D <= 4 ? 4 : 8);
    int Dp = (D + DC - 1) / DC * DC;
    return ((N + 63) / 64) * ((N + 127) / 128) * (Dp + 3);
}

int mll_bwd(stream_t, const MllBwdDesc& d) {
    if (d.nb <= 0 || d.nb % 128 != 0) return GPB_ERR_INVALID;
    if (!kind_valid(d.kind)) return GPB_ERR_INVALID;
    const int64_t N = d.N;
    const double var = d.variance[0];
    std::vector<double> acc(d.D, 0.0);
    double wk = 0.0, trw = 0.0, wshape = 0.0;
    for (int64_t r = 0; r < N; ++r)
        for (int64_t c = 0; c < N; ++c) {
            int64_t br = r / d.nb, bc = c / d.nb;
            if (br > bc) continue;
            double sv, wgt;
            if (br == bc) { sv = d.Sdiag[br * d.nb * d.nb + (r - br * d.nb) * d.nb + (c - br * d.nb)]; wgt = 1.0; }
            else { sv = d.S[r * d.lds + c]; wgt = 2.0; }
            double w = 0.5 * (d.alpha[r] * d.alpha[c] - sv);
            if (r == c) trw += w;
            w *= wgt;
            PairOut po;
            pair_eval(d.kind, d.X + r * d.ldx, d.X + c * d.ldx, d.D, d.ell, d.ell_is_scalar, d.variance, po);
            wk += w * po.k;
            wshape += w * po.dshape;
            for (int dd = 0; dd < d.D; ++dd) acc[dd] += w * po.dl[dd];
        }
    double g = d.gout ? d.gout[0] : 1.0;
    if (d.g_ell) {
        if (d.ell_is_scalar) {
            double s = 0.0;
            for (int dd = 0; dd < d.D; ++dd) s += g * acc[dd];
            d.g_ell[0] = s;
        } else
            for (int dd = 0; dd < d.D; ++dd) d.g_ell[dd] = g * acc[dd];
    }
    if (d.g_var) {
        d.g_var[0] = g * wk / var;
        if (kind_has_shape(d.kind)) d.g_var[1] = g * wshape;
    }
    if (d.g_obs_stddev) d.g_obs_stddev[0] = g * 2.0 * d.obs_stddev[0] * trw;
    if (d.g_mean) {
        double s = 0.0;
        for (int64_t i = 0; i < N; ++i) s += d.alpha[i];
        d.g_mean[0] = g * s;
    }
    return GPB_OK;
}

int64_t gram_bwd_partials_count(int64_t N, int64_t M, int D) {
    int DC = D <= 2 ? 2 : (D <= 4 ? 4 : 8);
    int Dp = (D + DC - 1) / DC * DC;
    return ((N + 63) / 64) * ((M + 127) / 128) * (Dp + 2);
}

int gram_bwd(stream_t, const GramBwdDesc& d) {
    if (!kind_valid(d.kind)) return GPB_ERR_INVALID;
    const double var = d.variance[0];
    std::vector<double> acc(d.D, 0.0);
    double wk = 0.0, wshape = 0.0;
    for (int64_t r = 0; r < d.N; ++r)
        for (int64_t c = 0; c < d.M; ++c) {
            double w = d.dK[r * d.lddk + c];
            PairOut po;
            pair_eval(d.kind, d.X + r * d.ldx, d.Z + c * d.ldz, d.D, d.ell, d.ell_is_scalar, d.variance, po);
            wk += w * po.k;
            wshape += w * po.dshape;
            for (int dd = 0; dd < d.D; ++dd) {
                acc[dd] += w * po.dl[dd];
                if (d.g_X) d.g_X[r * d.ldgx + dd] += d.scale * w * po.dx[dd];
                if (d.g_Z) d.g_Z[c * d.ldgz + dd] -= d.scale * w * po.dx[dd];
            }
        }
    if (d.g_ell) {
        if (d.ell_is_scalar) {
            for (int dd = 0; dd < d.D; ++dd) d.g_ell[0] += d.scale * acc[dd];
        } else
            for (int dd = 0; dd < d.D; ++dd) d.g_ell[dd] += d.scale * acc[dd];
    }
    if (d.g_var) {
        d.g_var[0] += d.scale * wk / var;
        if (kind_has_shape(d.kind)) d.g_var[1] += d.scale * wshape;
    }
    return GPB_OK;
}

}  // namespace gpb

// ---- SGPR helpers ---------------------------------------------------------------------------
namespace gpb {
int set_identity(stream_t, int64_t n, double* A, int64_t lda) {
    for (int64_t r = 0; r < n; ++r)
        for (int64_t c = 0; c < n; ++c) A[r * lda + c] = (r == c) ? 1.0 : 0.0;
    return GPB_OK;
}
int vec_sum(stream_t, int64_t n, const double* x, double* out) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += x[i];
    out[0] = s;
    return GPB_OK;
}
int axpy(stream_t, int64_t n, double alpha, const double* x, double* y) {
    for (int64_t i = 0; i < n; ++i) y[i] += alpha * x[i];
    return GPB_OK;
}
int scale_inplace(stream_t, int64_t n, double* x, const double* f) {
    if (!f || !x) return GPB_OK;
    for (int64_t i = 0; i < n; ++i) x[i] *= f[0];
    return GPB_OK;
}
int sgpr_aug_columns(stream_t, int64_t rows, double* T, int64_t ld, int64_t M, const double* y, const double* c) {
    for (int64_t r = 0; r < rows; ++r) {
        T[r * ld + M] = y[r] - (c ? c[0] : 0.0);
        T[r * ld + M + 1] = 1.0;
    }
    return GPB_OK;
}
int sgpr_prepare(stream_t, int64_t M, const double* P, int64_t ldp, const double* obs_stddev, double* Bmat,
                 double* psi, double* a1, double* sc) {
    double s = obs_stddev[0] * obs_stddev[0];
    double tr = 0.0;
    for (int64_t r = 0; r < M; ++r) {
        for (int64_t c = 0; c < M; ++c) {
            double v = (c <= r) ? P[r * ldp + c] : P[c * ldp + r];
            Bmat[r * M + c] = v / s + (r == c ? 1.0 : 0.0);
        }
        tr += P[r * ldp + r];
        psi[r] = P[M * ldp + r] / std::sqrt(s);
        a1[r] = P[(M + 1) * ldp + r] / std::sqrt(s);
    }
    sc[0] = P[M * ldp + M];
    sc[1] = P[(M + 1) * ldp + M];
    sc[2] = P[(M + 1) * ldp + M + 1];
    sc[3] = tr / s;
    sc[4] = s;
    return GPB_OK;
}
int sgpr_value(stream_t, const double* sc, const double* hl, const double* wtw, const double* variance,
               const int* info, double* out) {
    double dd = sc[0], n = sc[2], trphi = sc[3], s = sc[4];
    double v = 0.5 * ((-n * std::log(2.0 * M_PI * s) - 2.0 * hl[0] - (dd - wtw[0]) / s) - (n * variance[0] / s - trphi));
    if (info && (info[0] != 0 || info[1] != 0)) v = std::numeric_limits<double>::quiet_NaN();
    out[0] = v;
    return GPB_OK;
}
int sgpr_adjoints(stream_t, int64_t M, const double* Binv, const double* Bmat, const double* v, const double* sc,
                  double* G1, double* G2, double* u, double* rowsum) {
    double s = sc[4];
    for (int64_t r = 0; r < M; ++r) {
        double acc = 0.0;
        for (int64_t c = 0; c < M; ++c) {
            double eye = (r == c) ? 1.0 : 0.0;
            double dphi = 0.5 * (eye - Binv[r * M + c] - v[r] * v[c] / s);
            double phi = Bmat[r * M + c] - eye;
            G1[r * M + c] = (2.0 / s) * dphi;
            G2[r * M + c] = dphi - 0.5 * phi;
            acc += dphi * phi;
        }
        rowsum[r] = acc;
        u[r] = v[r] / (s * std::sqrt(s));
    }
    return GPB_OK;
}
int sgpr_scalar_grads(stream_t, const double* sc, const double* dots, const double* variance, const double* obs_stddev,
                      double* g_var, double* g_obs, double* g_mean) {
    double dd = sc[0], sd = sc[1], n = sc[2], s = sc[4];
    double psiv = dots[0], va1 = dots[1], dpp = dots[2];
    double g_s = -n / (2 * s) + (dd - psiv) / (2 * s * s) + n * variance[0] / (2 * s * s) - (2 * dpp + psiv / s) / (2 * s);
    if (g_var) g_var[0] += -n / (2 * s);
    if (g_obs) g_obs[0] = 2 * obs_stddev[0] * g_s;
    if (g_mean) g_mean[0] = -va1 / s + sd / s;
    return GPB_OK;
}
}  // namespace gpb

// ---- Ozaki (int8 digit plane) primitives: exact integer model of ozaki_i8.cu ------------------------------------
namespace gpb {
static int oz_row_exponent_host(double mx) {
    int e = std::ilogb(mx) + 2;
    if (std::scalbn(mx, -e) > 0.494) ++e;
    return e;
}
int ozaki_slice(stream_t, int64_t rows, int64_t k, int64_t kplane, const double* X, int64_t ldx, int nslices, int8_t* Q,
                int64_t ldq, double* scale, const int* nslices_dev) {
    if (rows < 0 || k <= 0 || kplane < k || nslices < 1 || nslices > OZ_PLANES_MAX || !X || !Q || !scale || ldq < (int64_t)nslices * kplane)
        return GPB_ERR_INVALID;
    if (nslices_dev) nslices = (nslices_dev[0] < 1 || nslices_dev[0] > nslices) ? nslices : nslices_dev[0];
    for (int64_t r = 0; r < rows; ++r) {
        double mx = 0.0;
        bool bad = false;
        for (int64_t c = 0; c < k; ++c) {
            double v = std::fabs(X[r * ldx + c]);
            if (!(v <= std::numeric_limits<double>::max())) bad = true;
            if (v > mx) mx = v;
        }
        int e = 0;
        if (bad) scale[r] = std::numeric_limits<double>::quiet_NaN();
        else if (mx == 0.0) scale[r] = 1.0;
        else { e = oz_row_exponent_host(mx); scale[r] = std::scalbn(1.0, e); }
        for (int64_t c = 0; c < kplane; ++c) {
            const double R = (bad || c >= k) ? 0.0 : std::scalbn(X[r * ldx + c], -e);
            long long I = std::llrint(std::scalbn(R, OZ_DIGIT_BITS * nslices));
            for (int p = nslices - 1; p >= 0; --p) {
                const int d = (int)((I + 128) & 255) - 128;
                Q[r * ldq + (int64_t)p * kplane + c] = (int8_t)d;
                I = (I - d) >> 8;
            }
            if (I != 0) return GPB_ERR_INVALID;  // the exponent rule keeps the top digit inside int8
        }
    }
    return GPB_OK;
}
int ozaki_slice_t(stream_t, int64_t rows, int64_t cols, int64_t kplane, const double* X, int64_t ldx, int nslices, int8_t* Qt,
                  int64_t ldq, double* scale, double* colmax_scratch) {
    if (rows <= 0 || cols <= 0 || kplane < rows || kplane % 128 || nslices < 1 || nslices > OZ_PLANES_MAX || !X || !Qt || !scale || !colmax_scratch ||
        ldq < (int64_t)nslices * kplane)
        return GPB_ERR_INVALID;
    for (int64_t c = 0; c < cols; ++c) {
        double mx = 0.0;
        bool bad = false;
        for (int64_t r = 0; r < rows; ++r) {
            double v = std::fabs(X[r * ldx + c]);
            if (!(v <= std::numeric_limits<double>::max())) bad = true;
            if (v > mx) mx = v;
        }
        colmax_scratch[c] = mx;
        int e = 0;
        if (bad) scale[c] = std::numeric_limits<double>::quiet_NaN();
        else if (mx == 0.0) scale[c] = 1.0;
        else { e = oz_row_exponent_host(mx); scale[c] = std::scalbn(1.0, e); }
        for (int64_t r = 0; r < kplane; ++r) {
            const double R = (bad || r >= rows) ? 0.0 : std::scalbn(X[r * ldx + c], -e);
            long long I = std::llrint(std::scalbn(R, OZ_DIGIT_BITS * nslices));
            for (int p = nslices - 1; p >= 0; --p) {
                const int d = (int)((I + 128) & 255) - 128;
                Qt[c * ldq + (int64_t)p * kplane + r] = (int8_t)d;
                I = (I - d) >> 8;
            }
            if (I != 0) return GPB_ERR_INVALID;
        }
    }
    return GPB_OK;
}
int col_weighted_sums(stream_t, int64_t rows, int64_t cols, const double* X, int64_t ldx, const double* w, double* scratch,
                      double* out_w, double* out_1) {
    if (rows <= 0 || cols <= 0 || !X || !w || !scratch || !out_w || !out_1) return GPB_ERR_INVALID;
    for (int64_t c = 0; c < cols; ++c) {
        double sw = 0.0, s1 = 0.0;
        for (int64_t r = 0; r < rows; ++r) { sw += w[r] * X[r * ldx + c]; s1 += X[r * ldx + c]; }
        out_w[c] += sw;
        out_1[c] += s1;
    }
    return GPB_OK;
}
int ozaki_gemm(stream_t, const OzakiGemmDesc& d) {
    if (d.M < 0 || d.N < 0 || d.K <= 0 || d.nslices < 1 || d.nslices > OZ_PLANES_MAX || !d.Qa || !d.Qb || !d.sa || !d.sb || !d.C)
        return GPB_ERR_INVALID;
    if (d.K % 128 || (int64_t)d.nslices * d.K * OZ_DIGIT_SQ_MAX >= (1ll << 31)) return GPB_ERR_UNSUPPORTED;
    if (d.mask == MASK_BLOCK_UPPER_DIAG_TO_C2 && (!d.C2 || d.mask_nb <= 0 || d.mask_nb % 128 || d.mask_col0 % 128)) return GPB_ERR_INVALID;
    const int planes = d.nslices_dev ? ((d.nslices_dev[0] < 1 || d.nslices_dev[0] > d.nslices) ? d.nslices : d.nslices_dev[0]) : d.nslices;
    const int64_t ps = d.plane_stride > 0 ? d.plane_stride : d.K;
    for (int64_t i = 0; i < d.M; ++i)
        for (int64_t j = 0; j < d.N; ++j) {
            const int64_t gr = d.mask_row0 + i, gc = d.mask_col0 + j;
            if (d.mask == MASK_LOWER && gr < gc) continue;
            if (d.mask == MASK_BLOCK_STRICT_UPPER && !(gr / d.mask_nb < gc / d.mask_nb)) continue;
            if (d.mask == MASK_BLOCK_UPPER_DIAG_TO_C2 && !(gr / d.mask_nb <= gc / d.mask_nb)) continue;
            // triangular operand: the kernel skips whole 128-digit K-blocks of structural zeros; the digits there ARE zero, so the
            // model may sum the full range -- but it must not read garbage, hence the same element-level range as GemmDesc::krange
            int64_t c0 = 0, c1 = d.K;
            if (d.krange == KR_B_LOWER) c1 = std::min<int64_t>(d.K, std::max<int64_t>(0, j + d.kr_off + 1));
            else if (d.krange == KR_B_UPPER) c0 = std::min<int64_t>(d.K, std::max<int64_t>(0, j + d.kr_off));
            else if (d.krange == KR_A_LOWER) c1 = std::min<int64_t>(d.K, std::max<int64_t>(0, i + d.kr_off + 1));
            else if (d.krange == KR_A_UPPER) c0 = std::min<int64_t>(d.K, std::max<int64_t>(0, i + d.kr_off));
            // exact int64 fixed-point recombination, as the default kernel (ozaki_i8.cu: oz_fx_bits / oz_store_row_fx)
            int lg = 0;
            while ((1ll << lg) < d.K * OZ_DIGIT_SQ_MAX) ++lg;
            int F = 61 - lg;
            if ((F & 7) == 7) --F;
            long long accq = 0;
            auto add_order = [&](int order, long long P) {
                const int sh = F - OZ_DIGIT_BITS * order;
                if (sh >= 0) accq += P * (1ll << sh);
                else accq += (P + (1ll << (-sh - 1))) >> (-sh);
            };
            for (int t = 0; t < planes; ++t) {
                int64_t P = 0;
                for (int p = 0; p <= t; ++p) {
                    const int8_t* a = d.Qa + i * d.ldqa + (int64_t)p * ps;
                    const int8_t* b = d.Qb + j * d.ldqb + (int64_t)(t - p) * ps;
                    int32_t part = 0;
                    for (int64_t c = c0; c < c1; ++c) part += (int32_t)a[c] * (int32_t)b[c];
                    P += part;
                }
                add_order(t, P);
            }
            if (planes >= 2 && planes % 2 == 0) {  // even plane count: the one order-`planes` pair of EQUAL planes (oz_has_diag)
                const int h = planes / 2;
                const int8_t* a = d.Qa + i * d.ldqa + (int64_t)h * ps;
                const int8_t* b = d.Qb + j * d.ldqb + (int64_t)h * ps;
                int32_t part = 0;
                for (int64_t c = c0; c < c1; ++c) part += (int32_t)a[c] * (int32_t)b[c];
                add_order(planes, part);
            }
            const double sr = d.alpha * d.sa[i] * std::scalbn(1.0, -F - 2 * OZ_DIGIT_BITS);
            const double v = (double)accq * (sr * d.sb[j]);
            double* dst = &d.C[i * d.ldc + j];
            if (d.mask == MASK_BLOCK_UPPER_DIAG_TO_C2 && gr / d.mask_nb == gc / d.mask_nb) {
                const int64_t b = gr / d.mask_nb;
                dst = d.C2 + b * d.mask_nb * d.mask_nb + (gr - b * d.mask_nb) * d.mask_nb + (gc - b * d.mask_nb);
            }
            *dst = d.beta0 ? v : *dst + v;
        }
    return GPB_OK;
}
int gram_digits(stream_t, const GramDigitsDesc& d) {
    const GramDesc& g = d.g;
    if (g.N <= 0 || g.M <= 0 || g.D <= 0 || d.nslices < 1 || d.nslices > OZ_PLANES_MAX) return GPB_ERR_INVALID;
    if (!g.X || !g.Z || !g.ell || !g.variance || !d.Q || !d.scale || !d.y) return GPB_ERR_INVALID;
    if (d.kplane % 128 || d.ldq < (int64_t)d.nslices * d.kplane) return GPB_ERR_INVALID;
    if (d.cols_mode ? (d.kplane < g.N || !d.part) : (d.kplane < g.M + 2)) return GPB_ERR_INVALID;
    if (g.kind == KIND_POWEXP || !kind_valid(g.kind)) return GPB_ERR_UNSUPPORTED;
    const double mean = d.mean_const ? d.mean_const[0] : 0.0;
    const double avar = std::fabs(g.variance[0]);
    const bool var_bad = !(avar <= std::numeric_limits<double>::max());
    const int e_var = (!var_bad && avar > 0.0) ? oz_row_exponent_host(avar) : 0;
    auto value = [&](int64_t r, int64_t c) -> double {
        if (r >= g.N) return 0.0;
        if (c < g.M) {
            PairOut po;
            pair_eval(g.kind, g.X + r * g.ldx, g.Z + c * g.ldz, g.D, g.ell, g.ell_is_scalar, g.variance, po);
            return po.k;
        }
        if (c == g.M) return d.y[r] - mean;
        if (c == g.M + 1) return 1.0;
        return 0.0;
    };
    auto digits = [&](double v, int e, bool bad, int8_t* out, int64_t stride) {
        long long I = bad ? 0 : std::llrint(std::scalbn(v, -e + OZ_DIGIT_BITS * d.nslices));
        for (int p = d.nslices - 1; p >= 0; --p) {
            const int dg = (int)((I + 128) & 255) - 128;
            out[(int64_t)p * stride] = (int8_t)dg;
            I = (I - dg) >> 8;
        }
        return I == 0;
    };
    if (!d.cols_mode) {
        for (int64_t r = 0; r < g.N; ++r) {
            const double dr = d.y[r] - mean;
            const double mx = std::max(std::max(avar, 1.0), std::fabs(dr));
            const bool bad = var_bad || !(mx <= std::numeric_limits<double>::max());
            const int e = bad ? 0 : oz_row_exponent_host(mx);
            d.scale[r] = bad ? std::numeric_limits<double>::quiet_NaN() : std::scalbn(1.0, e);
            for (int64_t c = 0; c < d.kplane; ++c)
                if (!digits(value(r, c), e, bad, d.Q + r * d.ldq + c, d.kplane)) return GPB_ERR_INVALID;
        }
        return GPB_OK;
    }
    const int64_t cols = g.M + 2, tiles = gram_digits_tile_rows(d.kplane);
    for (int64_t c = 0; c < g.M; ++c) {
        d.scale[c] = var_bad ? std::numeric_limits<double>::quiet_NaN() : (avar > 0.0 ? std::scalbn(1.0, e_var) : 1.0);
        for (int64_t r = 0; r < d.kplane; ++r)
            if (!digits(value(r, c), e_var, var_bad, d.Q + c * d.ldq + r, d.kplane)) return GPB_ERR_INVALID;
    }
    for (int64_t t = 0; t < tiles; ++t)
        for (int64_t c = 0; c < cols; ++c) {
            double sw = 0.0, s1 = 0.0;
            for (int64_t r = t * 64; r < std::min<int64_t>((t + 1) * 64, g.N); ++r) {
                const double v = value(r, c);
                sw += (d.y[r] - mean) * v;
                s1 += v;
            }
            d.part[(t * 2 + 0) * cols + c] = sw;
            d.part[(t * 2 + 1) * cols + c] = s1;
        }
    return GPB_OK;
}
int col_partials_reduce(stream_t, int64_t chunks, int64_t cols, const double* part, double* out_w, double* out_1) {
    if (chunks <= 0 || cols <= 0 || !part || !out_w || !out_1) return GPB_ERR_INVALID;
    for (int64_t c = 0; c < cols; ++c) {
        double sw = 0.0, s1 = 0.0;
        for (int64_t k = 0; k < chunks; ++k) { sw += part[(k * 2 + 0) * cols + c]; s1 += part[(k * 2 + 1) * cols + c]; }
        out_w[c] += sw;
        out_1[c] += s1;
    }
    return GPB_OK;
}
int igemm_i8(stream_t, int64_t m, int64_t n, int64_t k, const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* C,
             int64_t ldc) {
    for (int64_t i = 0; i < m; ++i)
        for (int64_t j = 0; j < n; ++j) {
            int32_t acc = 0;
            for (int64_t c = 0; c < k; ++c) acc += (int32_t)A[i * lda + c] * (int32_t)B[j * ldb + c];
            C[i * ldc + j] = acc;
        }
    return GPB_OK;
}
bool ozaki_available() { return true; }
int device_sm_count() { return 148; }
bool ozaki_supports_extensions() { return true; }
int ozaki_auto_planes_host(int64_t N, double variance, double obs_stddev, double jitter) {
    const double s = obs_stddev * obs_stddev + jitter;
    return ((double)N * std::fabs(variance) + s) / s <= OZ_AUTO_COND_LIMIT ? OZ_AUTO_PLANES_LO : OZ_AUTO_PLANES_HI;
}
int ozaki_choose_planes(stream_t, int requested, int64_t N, const double* variance, const double* obs_stddev, double jitter,
                        int* planes_out) {
    if (!planes_out || N < 0 || !(requested == OZ_AUTO || (requested >= 1 && requested <= OZ_PLANES_MAX))) return GPB_ERR_INVALID;
    if (requested != OZ_AUTO) planes_out[0] = requested;
    else planes_out[0] = (variance && obs_stddev) ? ozaki_auto_planes_host(N, variance[0], obs_stddev[0], jitter) : OZ_AUTO_PLANES_HI;
    return GPB_OK;
}
}  // namespace gpb

// measurement hooks are inert in the host model
namespace gpb {
stream_t side_stream(stream_t main, int) { return main; }
int stream_fork(stream_t, stream_t) { return GPB_OK; }
void profile_reset(int) {}
void debug_set_gemm_variant(int) {}
int profile_read(double* a, int64_t* b, int64_t* c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return GPB_OK; }
void profile_count_launch() {}
bool profile_enabled() { return false; }
void profile_gemm_begin(stream_t) {}
void profile_gemm_end(stream_t) {}
void profile_ozaki_begin(stream_t, double) {}
int profile_read_ozaki(double* a, int64_t* b, double* c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return GPB_OK; }
}  // namespace gpb

// ---- SVGP helpers -----------------------------------------------------------------------------
namespace gpb {
int sum_log_abs_diag(stream_t, int64_t n, const double* A, int64_t lda, double* out) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += std::log(std::fabs(A[i * (lda + 1)]));
    out[0] = s;
    return GPB_OK;
}
int svgp_unpack(stream_t, int64_t M, const double* P, int64_t ldp, const double* obs_stddev, double num_datapoints,
                double* Phi, double* psi, double* a1, double* sc) {
    double tr = 0.0;
    for (int64_t r = 0; r < M; ++r) {
        for (int64_t c = 0; c < M; ++c) Phi[r * M + c] = (c <= r) ? P[r * ldp + c] : P[c * ldp + r];
        tr += P[r * ldp + r];
        psi[r] = P[M * ldp + r];
        a1[r] = P[(M + 1) * ldp + r];
    }
    double s = obs_stddev[0] * obs_stddev[0], B = P[(M + 1) * ldp + M + 1];
    sc[0] = P[M * ldp + M]; sc[1] = P[(M + 1) * ldp + M]; sc[2] = B; sc[3] = tr; sc[4] = s;
    sc[5] = (num_datapoints / B) / s;
    return GPB_OK;
}
int svgp_value(stream_t, int64_t M, const double* sc, const double* dots, const double* variance, double jitter,
               const int* info, double* out) {
    double dd = sc[0], B = sc[2], trphi = sc[3], s = sc[4], coef = sc[5];
    double Q = dd - 2 * dots[0] + dots[3] + B * (variance[0] + jitter) - trphi;
    double ell = -0.5 * (B * std::log(2 * M_PI * s) + Q / s);
    double kl = 0.5 * (dots[1] - (double)M - 2 * dots[5] + 2 * dots[4] + dots[2]);
    double v = coef * s * ell - kl;
    if (info && info[0] != 0) v = std::numeric_limits<double>::quiet_NaN();
    out[0] = v;
    return GPB_OK;
}
int svgp_adjoints(stream_t, int64_t M, const double* Phi, const double* Tt, const double* PT, const double* u,
                  const double* psi, const double* sc, double* G1, double* E) {
    double coef = sc[5];
    for (int64_t r = 0; r < M; ++r)
        for (int64_t c = 0; c < M; ++c) {
            double eye = r == c ? 1.0 : 0.0, tt = Tt[r * M + c];
            G1[r * M + c] = coef * (eye - tt);
            E[r * M + c] = -0.5 * coef * (u[r] * psi[c] + u[c] * psi[r]) + 0.5 * coef * (PT[r * M + c] + PT[c * M + r]) -
                           0.5 * coef * Phi[r * M + c] + 0.5 * tt - 0.5 * eye;
        }
    return GPB_OK;
}
int svgp_vectors(stream_t, int64_t M, const double* psi, const double* Phiu, const double* u, const double* sc,
                 double* tvec, double* uvec) {
    for (int64_t i = 0; i < M; ++i) {
        tvec[i] = sc[5] * (psi[i] - Phiu[i]) - u[i];
        uvec[i] = sc[5] * u[i];
    }
    return GPB_OK;
}
int svgp_h(stream_t, int64_t M, const double* PhiV, const double* V, const double* sc, double* H) {
    for (int64_t i = 0; i < M * M; ++i) H[i] = sc[5] * PhiV[i] + V[i];
    return GPB_OK;
}
int svgp_gw_diag(stream_t, int64_t M, const double* W, int64_t ldw, double* gW, int64_t ldg) {
    for (int64_t i = 0; i < M; ++i) gW[i * ldg + i] += 1.0 / W[i * ldw + i];
    return GPB_OK;
}
int svgp_scalar_grads(stream_t, const double* sc, const double* dots, const double* dots2, const double* variance,
                      const double* obs_stddev, double jitter, double* g_var, double* g_obs, double* g_mean) {
    double dd = sc[0], sd = sc[1], B = sc[2], trphi = sc[3], s = sc[4], coef = sc[5], kappa = coef * s;
    double Q = dd - 2 * dots[0] + dots[3] + B * (variance[0] + jitter) - trphi;
    if (g_var) g_var[0] += -kappa * B / (2 * s);
    if (g_obs) g_obs[0] = 2 * obs_stddev[0] * kappa * (-B / (2 * s) + Q / (2 * s * s));
    if (g_mean) g_mean[0] = coef * (sd - dots2[0]) - dots2[1];
    return GPB_OK;
}
}  // namespace gpb
