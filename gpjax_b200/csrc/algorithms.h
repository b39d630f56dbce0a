// Blocked algorithms of the hot path, written purely against primitives.h (device pointers are
// opaque here: this file never dereferences them, so it runs unchanged on the CUDA primitives and
// on the host model used by the CPU tests).
#pragma once
#include <cstdint>
#include "primitives.h"

namespace gpb {

// Block size NB of every blocked algorithm (a multiple of 128): a function of the order the WORKSPACE was sized for, fixed when
// the workspace is carved (FactorWs::nb), so every call that shares a workspace -- the factorisation and the solves / inverse that
// reuse its diagonal-block inverses -- agrees on it.  Measured on B200 (scripts/nb_sweep.py, value + gradient of conjugate_mll):
//   round 1, FP64 DMMA path: N=50k  512: 3.93 s, 1024: 3.84 s, 2048: 3.82 s; N=20k 285 / 283 / 291 ms; N=10k 53.0 / 54.6 / 58.4 ms;
//   round 2, int8 path     : N=50k 1024: 1495 ms, 2048: 1305 ms, 4096: 1375 ms; N=20k 136.6 / 132.5 ms; N=16,384 88.3 / 85.2 ms;
//                            N=12,288 51.8 / 50.9 ms; N=10k 37.6 / 38.7 ms; N=8192 26.5 / 27.0 ms  (profiles/r02_nb_rule_sweeps.log)
// (K of every trailing update doubles with NB: half the passes over C and half the write-outs per int8 product -- the write-out is
// what paces that kernel -- against a longer latency-bound diagonal-block chain, which the look-ahead hides only for large N).
// -DGPB_NB=... (build.py: GPB_NB=...) forces one block size for every order: the host model of the CPU tests (256) and sweeps.
// GPB_NB_LARGE / GPB_NB_LARGE_MIN_ROWS (environment, read once): measurement hooks for the large-order rule (scripts/nb_sweep.py).
constexpr int64_t NB_SMALL = 1024, NB_LARGE = 2048, NB_LARGE_MIN_ROWS = 12288;
int64_t block_size_for(int64_t ws_n);
constexpr int64_t LEAFN = 128; // leaf size handled by potrf_leaf

inline int64_t nblocks(int64_t n, int64_t nb) { return (n + nb - 1) / nb; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- process-wide switch: arithmetic of the rank-NB trailing updates (>= OZ_MIN_ROWS output rows) and of the panel x inverse-
// diagonal-block products next to them (those also obey GPB_OZ_PANELS=0 -> DMMA; profiles/r02_cond_sweep_panels_{dmma,int8}.jsonl)
//   OZ_AUTO (-1, default): int8 digit planes on tcgen05 (balanced radix-256 digits, 8 bits per plane), plane count decided per
//                          call ON THE DEVICE by ozaki_choose_planes: the fused objectives use 6 planes (48 bits) only when the
//                          hyper-parameters bound cond(Sigma) by 2e6, else 7 (56 bits);
//                          a bare matrix (gpb_potrf_lower / gpb_potri_lower: nothing known about it) always gets 7;
//   4..7                 : that many planes everywhere (measurement / opt-in);
//   0                    : FP64 DMMA everywhere (also switches the SGPR / SVGP int8 products off).
// Why 7 unless proven benign (profiles/r02_cond_sweep_n8192.jsonl + r02_cond_sweep_radix256.jsonl, tests/test_gpu_conditioning.py,
// N = 8192, cond 1e3 .. 3e8): 56 bits below the row maximum are indistinguishable from the FP64 DMMA path against the CPU oracle
// at every conditioning, and the triangular solves agree element-wise to 2e-10; 48 bits (+ the equal-plane term) sit 10-300x
// above the DMMA path's error: at most 1.4e-15 x (the guard's bound) relative in the most sensitive gradient, ~3e-8 element-wise
// in the solves.
// Read from the environment variable GPB_OZAKI ("auto", 0, 4..7) at first use, overridable through set_ozaki_slices; the
// value is a std::atomic configuration word -- it is not meant to change while calls are in flight.
#ifndef GPB_OZ_DEFAULT
#define GPB_OZ_DEFAULT OZ_AUTO
#endif
void set_ozaki_slices(int nslices);
int get_ozaki_slices();
constexpr int OZ_MAX_SLICES = OZ_PLANES_MAX;  // 7 planes of 8 bits are extracted; the products use ws.oz_planes[0] of them
constexpr int OZ_MIN_SLICES = 4;
#ifndef GPB_OZ_MIN_ROWS
#define GPB_OZ_MIN_ROWS 2048
#endif
constexpr int64_t OZ_MIN_ROWS = GPB_OZ_MIN_ROWS;  // smaller updates stay on the DMMA pipe (launch + slicing overhead dominates)

// ---- workspace for the exact-GP factorisation family --------------------------------------
struct FactorWs {
    int64_t nb = 0;           // block size NB of every call on this workspace: block_size_for(the order it was carved for)
    double* Dinv = nullptr;   // nblk blocks [NB x NB], inverse of the diagonal blocks of L
    double* DinvT = nullptr;  // their transposes
    double* Sdiag = nullptr;  // nblk blocks [NB x NB], diagonal blocks of Sigma^-1 (potri only)
    double* panel = nullptr;  // [N x NB] contiguous panel copy
    double* panel2 = nullptr; // second panel buffer (lookahead double-buffering)
    double* small = nullptr;  // 4 x [NB x NB] scratch
    double* vec = nullptr;    // 4 x [N] vectors (d, w, tmp, spare)
    double* scal = nullptr;   // 16 scalars
    double* partials = nullptr;
    int64_t partials_count = 0;
    // Ozaki path (null when N < OZ_MIN_ROWS + NB): two digit buffers [N x OZ_MAX_SLICES*NB] int8 and two row-scale vectors
    // (double-buffered like panel / panel2 so the lookahead may slice panel k+1 while update k still reads panel k)
    int8_t* oz_q = nullptr;
    int8_t* oz_q2 = nullptr;
    double* oz_scale = nullptr;
    double* oz_scale2 = nullptr;
    int* oz_planes = nullptr;  // device word: planes the int8 updates of the current call use (ozaki_choose_planes)
};
// Writes ws.oz_planes for the next factorisation-family call on this workspace: `variance` / `obs_stddev` (device scalars) of
// the covariance being factored when known (fused objectives), nullptr for a bare matrix.  No-op on the DMMA path.
int factor_set_planes(stream_t s, const FactorWs& ws, int64_t N, const double* variance, const double* obs_stddev, double jitter);
// bytes needed for an N x N problem with D input dims (with_potri: include Sdiag + backward partials)
int64_t factor_ws_bytes(int64_t N, int D, int with_potri);
// carve `ws` out of a caller buffer; returns GPB_ERR_WORKSPACE if too small
int factor_ws_carve(void* buf, int64_t bytes, int64_t N, int D, int with_potri, FactorWs* ws);

// ---- factorisation building blocks ----------------------------------------------------------
// In-place lower Cholesky of the lower triangle of A (strict upper never touched).  Fills
// ws.Dinv / ws.DinvT.  *info (device) receives the first failing pivot (1-based) or stays 0.
int potrf_lower(stream_t s, int64_t N, double* A, int64_t lda, const FactorWs& ws, int* info);
// Only invert the diagonal blocks of an existing lower-triangular L into ws.Dinv / ws.DinvT.
int diag_inverses(stream_t s, int64_t N, const double* L, int64_t lda, const FactorWs& ws);
// x <- L^-1 x   /   x <- L^-T x   (vector, in place), using ws.Dinv(T) and ws.vec[2N..3N) as scratch
int trsv_lower(stream_t s, int64_t N, const double* L, int64_t lda, const FactorWs& ws, double* x, int trans);
// B <- L^-1 B  /  B <- L^-T B  for an N x T row-major block B (in place); scratch: ws.panel must hold
// NB x T doubles (caller guarantees N*NB >= NB*T or provides a bigger panel)
int trsm_lower_left(stream_t s, int64_t N, int64_t T, const double* L, int64_t lda, const FactorWs& ws, double* B,
                    int64_t ldb, int trans);
// W = L^-T written into the blocks strictly above the block diagonal of A (diagonal blocks of W are
// ws.DinvT); L (lower, incl. diagonal blocks) is left intact.
int trtri_into_upper(stream_t s, int64_t N, double* A, int64_t lda, const FactorWs& ws);
// Sigma^-1 = W W^T: strictly-upper blocks overwrite W in A, diagonal blocks go to ws.Sdiag.
int lauum_upper(stream_t s, int64_t N, double* A, int64_t lda, const FactorWs& ws);

// ---- exact-GP objective (gpjax/objectives.py:93-107 + gpjax/distributions.py:124-134) --------
struct MllArgs {
    int kind = 0;
    int64_t N = 0;
    int D = 0;
    const double* X = nullptr;
    int64_t ldx = 0;
    const double* y = nullptr;           // [N]
    const double* ell = nullptr;         // device [D] or [1]
    int ell_is_scalar = 0;
    const double* variance = nullptr;    // device scalar
    const double* obs_stddev = nullptr;  // device scalar
    const double* mean_const = nullptr;  // device scalar or null (Zero mean)
    double jitter = 1e-6;
    double* Sigma = nullptr;  // N x N scratch/result buffer (lower: L; upper blocks: Sigma^-1 after bwd)
    int64_t lds = 0;
};
// value_out[0] = log N(y | m, K + (jitter + obs_stddev^2) I); alpha_out[N] = Sigma^-1 (y - m)
int mll_forward(stream_t s, const MllArgs& a, const FactorWs& ws, double* value_out, double* alpha_out, int* info);
// gradients w.r.t. (lengthscale, variance, obs_stddev, mean_const), scaled by *gout (null -> 1).
// Needs the Sigma buffer + ws exactly as mll_forward left them.
int mll_backward(stream_t s, const MllArgs& a, const FactorWs& ws, const double* alpha, const double* gout,
                 double* g_ell, double* g_var, double* g_obs_stddev, double* g_mean);

}  // namespace gpb
