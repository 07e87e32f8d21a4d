// Post-sweep kernels: Geweke_Z stop rule (R/clone_id.R:688-715), posterior means (:645-653),
// prob_variant (:606-622), the general get_logLik (:734-752) and clone_id_EM (:240-326).
#include "internal.h"

namespace cdl {

namespace {

// One thread per trace column.  X is [n_rows][cols]; windows A = rows [0, floor(.1 n)),
// B = rows [ceil(.5 n) - 1, n); both variances divide by nrow(A) - 1 as the reference does.
__global__ void k_geweke(const double* __restrict__ X, int64_t cols, int64_t n, unsigned long long* counts) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned ok = 0, conv = 0;
    if (c < cols) {
        const int64_t nA = (int64_t)floor(0.1 * (double)n);
        const int64_t bs = (int64_t)ceil(0.5 * (double)n) - 1;
        const int64_t nB = n - bs;
        double sA = 0.0, sB = 0.0;
        for (int64_t r = 0; r < nA; ++r) sA += X[r * cols + c];
        for (int64_t r = bs; r < n; ++r) sB += X[r * cols + c];
        const double mA = sA / (double)nA, mB = sB / (double)nB;
        double vA = 0.0, vB = 0.0;
        for (int64_t r = 0; r < nA; ++r) { const double d = X[r * cols + c] - mA; vA += d * d; }
        for (int64_t r = bs; r < n; ++r) { const double d = X[r * cols + c] - mB; vB += d * d; }
        vA /= (double)(nA - 1);
        vB /= (double)(nA - 1);
        const double Z = (mA - mB) / sqrt(vA + vB + 1e-50);
        if (!isnan(Z)) { ok = 1; conv = fabs(Z) <= 2.0; }
    }
    const unsigned okw = __popc(__ballot_sync(0xffffffffu, ok));
    const unsigned cvw = __popc(__ballot_sync(0xffffffffu, conv));
    if ((threadIdx.x & 31) == 0) {
        if (okw) atomicAdd(&counts[0], (unsigned long long)okw);
        if (cvw) atomicAdd(&counts[1], (unsigned long long)cvw);
    }
}

// The same test for SMALL traces (a few thousand columns: the C1 / C2 / C5 donors).  One thread per column leaves such
// a trace on 7 CTAs with ~2000 dependent trips per pass -- 0.5 ms per check, and the 64 chains x 30 checks of a
// clone_id(n_chain = 64) call on example_donor spent 0.65 of their 1.1 s there.  Here a CTA owns GW_TILE columns and
// deals the rows of each window to GW_TILE row lanes (partials added in lane order: deterministic; the sums differ from
// the one-thread walk by rounding only), and blockIdx.y walks the traces of several chains in one launch
// (counts[2 y], counts[2 y + 1]).
constexpr int GW_TILE = 32;
__global__ void __launch_bounds__(GW_TILE * GW_TILE)
k_geweke_tiled(const double* __restrict__ X0, const double* const* __restrict__ Xs, int64_t cols, int64_t n,
               unsigned long long* counts) {
    __shared__ double red[2][GW_TILE][GW_TILE + 1];
    __shared__ double mean[2][GW_TILE];
    const double* __restrict__ X = Xs ? Xs[blockIdx.y] : X0;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t c = blockIdx.x * (int64_t)GW_TILE + tx;
    const bool in = c < cols;
    const int64_t nA = (int64_t)floor(0.1 * (double)n);
    const int64_t bs = (int64_t)ceil(0.5 * (double)n) - 1;
    const int64_t nB = n - bs;
    double sA = 0.0, sB = 0.0;
    if (in) {
#pragma unroll 4
        for (int64_t r = ty; r < nA; r += GW_TILE) sA += X[r * cols + c];
#pragma unroll 4
        for (int64_t r = bs + ty; r < n; r += GW_TILE) sB += X[r * cols + c];
    }
    red[0][ty][tx] = sA; red[1][ty][tx] = sB;
    __syncthreads();
    if (ty == 0) {
        double a = 0.0, b = 0.0;
        for (int q = 0; q < GW_TILE; ++q) { a += red[0][q][tx]; b += red[1][q][tx]; }
        mean[0][tx] = a / (double)nA; mean[1][tx] = b / (double)nB;
    }
    __syncthreads();
    const double mA = mean[0][tx], mB = mean[1][tx];
    double vA = 0.0, vB = 0.0;
    if (in) {
#pragma unroll 4
        for (int64_t r = ty; r < nA; r += GW_TILE) { const double d = X[r * cols + c] - mA; vA += d * d; }
#pragma unroll 4
        for (int64_t r = bs + ty; r < n; r += GW_TILE) { const double d = X[r * cols + c] - mB; vB += d * d; }
    }
    red[0][ty][tx] = vA; red[1][ty][tx] = vB;       // (the partial sums were read before the barrier above)
    __syncthreads();
    if (ty == 0) {                                  // warp 0: one lane per column of the tile
        double a = 0.0, b = 0.0;
        for (int q = 0; q < GW_TILE; ++q) { a += red[0][q][tx]; b += red[1][q][tx]; }
        a /= (double)(nA - 1);
        b /= (double)(nA - 1);                      // nrow(A) - 1, as the reference has it
        const double Z = (mA - mB) / sqrt(a + b + 1e-50);
        unsigned ok = 0, conv = 0;
        if (in && !isnan(Z)) { ok = 1; conv = fabs(Z) <= 2.0; }
        const unsigned okw = __popc(__ballot_sync(0xffffffffu, ok));
        const unsigned cvw = __popc(__ballot_sync(0xffffffffu, conv));
        if (tx == 0) {
            if (okw) atomicAdd(&counts[2 * blockIdx.y], (unsigned long long)okw);
            if (cvw) atomicAdd(&counts[2 * blockIdx.y + 1], (unsigned long long)cvw);
        }
    }
}
// traces up to this many columns take the tiled kernels (beyond it one thread per column fills the GPU by itself)
constexpr int64_t TILED_MAX_COLS = 1 << 16;

template <typename T>
__global__ void __launch_bounds__(GW_TILE * GW_TILE)
k_col_means_tiled(const T* __restrict__ X, int64_t cols, int64_t r0, int64_t r1, double* __restrict__ out) {
    __shared__ double red[GW_TILE][GW_TILE + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t c = blockIdx.x * (int64_t)GW_TILE + tx;
    double s = 0.0;
    if (c < cols) {
#pragma unroll 4
        for (int64_t r = r0 + ty; r < r1; r += GW_TILE) s += (double)X[r * cols + c];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
        double a = 0.0;
        for (int q = 0; q < GW_TILE; ++q) a += red[q][tx];
        out[c] = a / (double)(r1 - r0);
    }
}

template <typename T>
__global__ void k_col_means(const T* __restrict__ X, int64_t cols, int64_t r0, int64_t r1, double* __restrict__ out) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += (double)X[r * cols + c];
    out[c] = s / (double)(r1 - r0);
}

// prob_variant for wise in {global, variant}: per covered entry
//   sum_t dbinom(a, d, th1_t) / (sum_t dbinom(a, d, th1_t) + sum_t dbinom(a, d, th0_t)); the choose(d, a)
// factor is common and cancels.  Uncovered entries are 0.5 (dbinom(0, 0, .) = 1), filled by the caller.
template <typename VIdx>
__global__ void k_prob_variant(const uint32_t* __restrict__ rowgrp, const VIdx* __restrict__ vidx,
                               const uint16_t* __restrict__ ca, const uint16_t* __restrict__ cb, int64_t M,
                               int64_t N, int wise, const double* __restrict__ th0, const double* __restrict__ th1,
                               int64_t n_el, int64_t r0, int64_t r1, const int64_t* __restrict__ elem0,
                               double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < M; j += nwarps) {
        const int64_t e0 = (int64_t)rowgrp[j] * GROUP, e1 = (int64_t)rowgrp[j + 1] * GROUP;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const uint32_t a = ca[e], b = cb[e];
            if (a + b == 0) continue;
            const int64_t i = vidx[e];
            // sum_t dbinom(a, d, th1_t) / (sum_t dbinom(a, d, th1_t) + sum_t dbinom(a, d, th0_t)): choose(d, a) is common
            // to every term and cancels; so does exp(mx), mx the largest exponent of either arm -- without it deep
            // entries (mtDNA: a = 326 of d = 6213 has exponents near -1300) underflow to 0 / 0, where R's dbinom
            // stays representable because it carries the binomial coefficient.
            const int64_t t1i = wise == 2 ? elem0[j] + (e - e0) : (wise == 1 ? i : 0);
            double mx = -INFINITY;
            for (int64_t r = r0; r < r1; ++r) {
                const double t1 = th1[r * n_el + t1i], t0 = th0[r];
                mx = fmax(mx, fmax((double)a * log(t1) + (double)b * log(1.0 - t1),
                                   (double)a * log(t0) + (double)b * log(1.0 - t0)));
            }
            double p1 = 0.0, p0 = 0.0;
            for (int64_t r = r0; r < r1; ++r) {
                const double t1 = th1[r * n_el + t1i], t0 = th0[r];
                p1 += exp((double)a * log(t1) + (double)b * log(1.0 - t1) - mx);
                p0 += exp((double)a * log(t0) + (double)b * log(1.0 - t0) - mx);
            }
            out[i + j * N] = p1 / (p1 + p0);
        }
    }
}

__global__ void k_fill(double* p, int64_t n, double v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// get_logLik with arbitrary (fractional) Config and either labels or an assignment-probability matrix.
// Dense definition per covered entry: log( exp(l1 a + l1c b) * q + exp(l0 a + l0c b) * (1 - q) ),
// q = sum_k Config[i,k] * P[j,k]; uncovered entries contribute log(1) = 0.  One warp per cell, partial
// sums per cell then a fixed-order reduction (deterministic).
template <typename VIdx>
__global__ void k_loglik(const uint32_t* __restrict__ rowgrp, const VIdx* __restrict__ vidx,
                         const uint16_t* __restrict__ ca, const uint16_t* __restrict__ cb, int64_t M, int64_t N,
                         int K, const double* __restrict__ Config /* N x K col-major */,
                         const int32_t* __restrict__ assign, const double* __restrict__ prob /* [M][K] */,
                         double l0, double l0c, const double* __restrict__ theta1, int64_t n_th1,
                         const int64_t* __restrict__ elem0 /* per-entry theta1 (n_th1 == nnz) or nullptr */,
                         double* __restrict__ cell_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < M; j += nwarps) {
        const int64_t e0 = (int64_t)rowgrp[j] * GROUP, e1 = (int64_t)rowgrp[j + 1] * GROUP;
        const int zlab = assign ? assign[j] - 1 : -1;
        double s = 0.0;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const uint32_t a = ca[e], b = cb[e];
            if (a + b == 0) continue;
            const int64_t i = vidx[e];
            double q = 0.0;
            if (assign) q = Config[i + (int64_t)zlab * N];
            else
                for (int k = 0; k < K; ++k) q += Config[i + (int64_t)k * N] * prob[j * K + k];
            const double t1 = theta1[elem0 ? elem0[j] + (e - e0) : (n_th1 == 1 ? 0 : i)];
            const double w1 = exp(log(t1) * (double)a + log(1.0 - t1) * (double)b);
            const double w0 = exp(l0 * (double)a + l0c * (double)b);
            s += log(w1 * q + w0 * (1.0 - q));
        }
        s = warp_sum_d(s);
        if (lane == 0) cell_out[j] = s;
    }
}

// fixed-order sum of n doubles by one block (deterministic)
__global__ void k_sum_fixed(const double* __restrict__ x, int64_t n, double* out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// ---- EM ----
// S3[j,k] = sum_i A_ij C_ik, S4 likewise with B (R/clone_id.R:274-275); S1 = TA - S3, S2 = TB - S4.
template <typename VIdx>
__global__ void k_em_stats(const uint32_t* __restrict__ rowgrp, const VIdx* __restrict__ vidx,
                           const uint16_t* __restrict__ ca, const uint16_t* __restrict__ cb, int64_t M, int K,
                           const uint32_t* __restrict__ cfg, uint32_t* __restrict__ S3, uint32_t* __restrict__ S4,
                           uint32_t* __restrict__ TA, uint32_t* __restrict__ TB) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < M; j += nwarps) {
        const int64_t e0 = (int64_t)rowgrp[j] * GROUP, e1 = (int64_t)rowgrp[j + 1] * GROUP;
        uint32_t s3[KMAX], s4[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) { s3[k] = 0; s4[k] = 0; }
        uint32_t ta = 0, tb = 0;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const uint32_t a = ca[e], b = cb[e];
            const uint32_t m = cfg[vidx[e]];
            ta += a; tb += b;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                const uint32_t on = 0u - ((m >> k) & 1u);
                s3[k] += a & on;
                s4[k] += b & on;
            }
        }
        ta = __reduce_add_sync(0xffffffffu, ta);
        tb = __reduce_add_sync(0xffffffffu, tb);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                const uint32_t x = __reduce_add_sync(0xffffffffu, s3[k]);
                const uint32_t y = __reduce_add_sync(0xffffffffu, s4[k]);
                if (lane == 0) { S3[j * K + k] = x; S4[j * K + k] = y; }
            }
        }
        if (lane == 0) { TA[j] = ta; TB[j] = tb; }
    }
}

// One EM iteration over the M x K statistics (R/clone_id.R:291-317).  Per-block partial sums of
// (logSumExp, prob*S1, prob*(S1+S2), prob*S3, prob*(S3+S4)) go to `partials`; k_em_final adds them in
// block order.
__global__ void k_em_iter(int64_t M, int K, const uint32_t* __restrict__ S3, const uint32_t* __restrict__ S4,
                          const uint32_t* __restrict__ TA, const uint32_t* __restrict__ TB,
                          const double* __restrict__ logPsi, double l0, double l0c, double l1, double l1c,
                          double* __restrict__ prob, double* __restrict__ partials) {
    __shared__ double sm[5][8];
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double v[5] = {0, 0, 0, 0, 0};
    if (j < M) {
        double L[KMAX];
        double mx = -INFINITY;
        const double ta = TA[j], tb = TB[j];
        for (int k = 0; k < K; ++k) {
            const double s3 = S3[j * K + k], s4 = S4[j * K + k];
            L[k] = (ta - s3) * l0 + (tb - s4) * l0c + s3 * l1 + s4 * l1c + logPsi[k];
            mx = fmax(mx, L[k]);
        }
        double se = 0.0;
        for (int k = 0; k < K; ++k) { L[k] = exp(L[k] - mx); se += L[k]; }
        v[0] = mx + log(se);
        for (int k = 0; k < K; ++k) {
            const double p = L[k] / se;
            prob[j * K + k] = p;
            const double s3 = S3[j * K + k], s4 = S4[j * K + k];
            const double s1 = ta - s3, s2 = tb - s4;
            v[1] += p * s1; v[2] += p * (s1 + s2); v[3] += p * s3; v[4] += p * (s3 + s4);
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double r = warp_sum_d(v[q]);
        if (lane == 0) sm[q][w] = r;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double r = 0.0;
        for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) r += sm[threadIdx.x][ww];
        partials[(size_t)blockIdx.x * 5 + threadIdx.x] = r;
    }
}

__global__ void k_em_final(const double* __restrict__ partials, int64_t nblocks, double* __restrict__ out5) {
    __shared__ double sm[256];
    for (int q = 0; q < 5; ++q) {
        double s = 0.0;
        for (int64_t b = threadIdx.x; b < nblocks; b += blockDim.x) s += partials[b * 5 + q];
        sm[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out5[q] = sm[0];
        __syncthreads();
    }
}

inline unsigned warp_grid(int64_t M, int threads) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>((M * 32 + threads - 1) / threads, 148 * 16));
}

}  // namespace

namespace {
// out[c * R + r] = in[r * C + c] (row-major R x C -> column-major R x C, i.e. R's layout), optionally scaled
template <bool SWAP>
__global__ void k_transpose_scale(const double* __restrict__ in, int64_t R, int64_t C, double scale,
                                  double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)(SWAP ? blockIdx.y : blockIdx.x) * 32, c0 = (int64_t)(SWAP ? blockIdx.x : blockIdx.y) * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int64_t r = r0 + dy, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[dy][threadIdx.x] = in[r * C + c] * scale;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int64_t c = c0 + dy, r = r0 + threadIdx.x;
        if (r < R && c < C) out[c * R + r] = tile[threadIdx.x][dy];
    }
}

// The same Z test (R/clone_id.R:688-715) from RUNNING column sums instead of the stored trace: s1 / s2 are the
// sums of x and x^2 over rows 1..r, snapshotted at r = a = floor(.1 n) (window A), r = b = ceil(.5 n) - 1 (the
// rows before window B) and current at r = n.  A null snapshot stands for r <= 1 (row 1 of prob_all is all zero).
// One-pass variances (clamped at 0) differ from the two-pass ones of k_geweke by rounding only, relative to the
// column's mean square.
template <typename S>
__global__ void k_geweke_sums(const S* __restrict__ s1a, const S* __restrict__ s2a,
                              const S* __restrict__ s1b, const S* __restrict__ s2b,
                              const double* __restrict__ s1n, const double* __restrict__ s2n, int64_t cols, int64_t n,
                              unsigned long long* counts) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned ok = 0, conv = 0;
    if (c < cols) {
        const double nA = floor(0.1 * (double)n);
        const double nB = (double)n - (ceil(0.5 * (double)n) - 1.0);
        const double a1 = s1a ? (double)s1a[c] : 0.0, a2 = s2a ? (double)s2a[c] : 0.0;
        const double b1 = s1n[c] - (s1b ? (double)s1b[c] : 0.0), b2 = s2n[c] - (s2b ? (double)s2b[c] : 0.0);
        const double mA = a1 / nA, mB = b1 / nB;
        // One-pass variances.  Their cancellation error is ~1e-16 (fp64 snapshots) or ~6e-8 (fp32) of the column's
        // mean square m^2 -- and never decides Z: window A contains trace row 1, which is all zero (the reference
        // starts filling prob_all at row 2), so vA >= ~m^2 / nA for every column and dominates that error.
        const double vA = fmax(a2 - nA * mA * mA, 0.0) / (nA - 1.0);
        const double vB = fmax(b2 - nB * mB * mB, 0.0) / (nA - 1.0);     // nrow(A) - 1, as the reference has it
        const double Z = (mA - mB) / sqrt(vA + vB + 1e-50);
        if (!isnan(Z)) { ok = 1; conv = fabs(Z) <= 2.0; }
    }
    const unsigned okw = __popc(__ballot_sync(0xffffffffu, ok));
    const unsigned cvw = __popc(__ballot_sync(0xffffffffu, conv));
    if ((threadIdx.x & 31) == 0) {
        if (okw) atomicAdd(&counts[0], (unsigned long long)okw);
        if (cvw) atomicAdd(&counts[1], (unsigned long long)cvw);
    }
}


// out = (s1n - s1u) * inv: column means over the rows after snapshot u
template <typename S>
__global__ void k_window_mean(const double* __restrict__ s1n, const S* __restrict__ s1u, int64_t n, double inv,
                              double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (s1n[i] - (s1u ? (double)s1u[i] : 0.0)) * inv;
}

__global__ void k_scale(double* __restrict__ x, int64_t n, double s) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] *= s;
}


// ---- relabel (R/clone_id.R:624-644) on the device ---------------------------------------------------------------
// For every trace row ii from n_buin on, the clone columns of prob_all[ii] are matched to those of prob_all[ii - 1]
// (already relabelled) by colMatch (R/assessment.R:87-121): minimum mean absolute difference over all K!
// permutations for K <= 5, greedy above; the reference then COMPOSES the match into col_idx_use and permutes the
// row with col_idx_use (not with the match itself) -- kept as written.  Rows depend on their predecessor, so the
// rows are walked in order: cost matrix (one pass over the two rows), solve + compose (one thread), permute.
constexpr int RELABEL_TILE = 64;
__global__ void __launch_bounds__(256) k_relabel_cost(const double* __restrict__ A, const double* __restrict__ B,
                                                      int64_t M, int K, double* __restrict__ partials) {
    __shared__ double tA[RELABEL_TILE * KMAX], tB[RELABEL_TILE * KMAX];
    const int KK = K * K;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};            // pairs p = tid + 256 q, q < 4 (K*K <= 1024)
    for (int64_t c0 = (int64_t)blockIdx.x * RELABEL_TILE; c0 < M; c0 += (int64_t)gridDim.x * RELABEL_TILE) {
        const int nc = (int)min((int64_t)RELABEL_TILE, M - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < nc * K; t += blockDim.x) { tA[t] = A[c0 * K + t]; tB[t] = B[c0 * K + t]; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = threadIdx.x + 256 * q;
            if (p < KK) {
                const int i = p / K, j = p - i * K;
                double s = 0.0;
                for (int c = 0; c < nc; ++c) s += fabs(tA[c * K + i] - tB[c * K + j]);
                acc[q] += s;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int p = threadIdx.x + 256 * q;
        if (p < KK) partials[(size_t)blockIdx.x * KK + p] = acc[q];
    }
}

__global__ void __launch_bounds__(256) k_relabel_solve(const double* __restrict__ partials, int nblocks, int K, int64_t M,
                                                       int* __restrict__ col_idx_use) {
    __shared__ double cost[KMAX * KMAX];
    const int KK = K * K;
    for (int p = threadIdx.x; p < KK; p += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * KK + p];     // blocks in order: fixed rounding
        cost[p] = s / (double)M;                                                 // MAE_mat[i, j]
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int idx[KMAX];
    if (K <= 5) {
        // all permutations in lexicographic order, first minimum (which.min)
        int perm[5], best[5];
        for (int k = 0; k < K; ++k) perm[k] = k;
        double bv = INFINITY;
        for (;;) {
            double v = 0.0;
            for (int i = 0; i < K; ++i) v += cost[i * K + perm[i]];
            if (v < bv) { bv = v; for (int i = 0; i < K; ++i) best[i] = perm[i]; }
            int i = K - 2;
            while (i >= 0 && perm[i] > perm[i + 1]) --i;
            if (i < 0) break;
            int j = K - 1;
            while (perm[j] < perm[i]) --j;
            int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
            for (int l = i + 1, r = K - 1; l < r; ++l, --r) { t = perm[l]; perm[l] = perm[r]; perm[r] = t; }
        }
        for (int i = 0; i < K; ++i) idx[i] = best[i];
    } else {
        // greedy: repeatedly the smallest remaining entry (first in column-major order), then its row and column
        // are struck out
        double big = -INFINITY;
        for (int p = 0; p < KK; ++p) big = fmax(big, cost[p]);
        big += 1.0;
        for (int i = 0; i < K; ++i) idx[i] = -1;
        for (int step = 0; step < K; ++step) {
            double mv = INFINITY; int mr = 0, mc = 0;
            for (int c = 0; c < K; ++c)
                for (int r = 0; r < K; ++r)
                    if (cost[r * K + c] < mv) { mv = cost[r * K + c]; mr = r; mc = c; }
            idx[mr] = mc;
            for (int q = 0; q < K; ++q) { cost[mr * K + q] = big; cost[q * K + mc] = big; }
        }
    }
    int nu[KMAX];
    for (int i = 0; i < K; ++i) nu[i] = col_idx_use[idx[i]];                     // col_idx_use <- col_idx_use[idx]
    for (int i = 0; i < K; ++i) col_idx_use[i] = nu[i];
}

__global__ void k_relabel_apply(double* __restrict__ prob_row, uint8_t* __restrict__ cfg_row, int64_t M, int64_t N, int K,
                                const int* __restrict__ col_idx_use) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < M) {
        double v[KMAX];
        for (int k = 0; k < K; ++k) v[k] = prob_row[t * K + col_idx_use[k]];
        for (int k = 0; k < K; ++k) prob_row[t * K + k] = v[k];
    }
    if (t < N) {
        uint8_t v[KMAX];
        for (int k = 0; k < K; ++k) v[k] = cfg_row[t + (int64_t)col_idx_use[k] * N];
        for (int k = 0; k < K; ++k) cfg_row[t + (int64_t)k * N] = v[k];
    }
}

}  // namespace

void relabel_trace(double* prob_all, uint8_t* config_all, int64_t M, int64_t N, int K, int64_t n_buin, int64_t it,
                   cudaStream_t st) {
    if (n_buin < 2 || it < n_buin) return;       // row n_buin needs its predecessor
    const int nblocks = (int)std::max<int64_t>(1, std::min<int64_t>((M + RELABEL_TILE - 1) / RELABEL_TILE, 296));
    double* partials = nullptr;
    int* use = nullptr;
    CDL_CUDA(cudaMalloc(&partials, sizeof(double) * (size_t)nblocks * K * K));
    cudaError_t e = cudaMalloc(&use, sizeof(int) * KMAX);
    if (e != cudaSuccess) { cudaFree(partials); CDL_CUDA(e); }
    int ident[KMAX];
    for (int k = 0; k < KMAX; ++k) ident[k] = k;
    cudaMemcpyAsync(use, ident, sizeof(ident), cudaMemcpyHostToDevice, st);
    const int64_t mx = std::max(M, N);
    for (int64_t ii = n_buin; ii <= it; ++ii) {
        double* prev = prob_all + (ii - 2) * M * K;
        double* cur = prob_all + (ii - 1) * M * K;
        k_relabel_cost<<<nblocks, 256, 0, st>>>(prev, cur, M, K, partials);
        k_relabel_solve<<<1, 256, 0, st>>>(partials, nblocks, K, M, use);
        k_relabel_apply<<<(unsigned)((mx + 255) / 256), 256, 0, st>>>(cur, config_all + (ii - 1) * N * K, M, N, K, use);
    }
    e = cudaStreamSynchronize(st);
    cudaFree(partials);
    cudaFree(use);
    CDL_CUDA(e);
    CDL_CUDA(cudaGetLastError());
}

namespace {
}  // namespace

void launch_transpose_scale(const double* in, int64_t R, int64_t C, double scale, double* out, cudaStream_t st) {
    if (R <= 0 || C <= 0) return;
    const int64_t rb = (R + 31) / 32, cbk = (C + 31) / 32;      // the longer side goes on grid.x (no 65535 limit)
    dim3 block(32, 8);
    if (rb >= cbk) {
        if (cbk > 65535) throw Error(5, "matrix too large for the transpose kernel");
        k_transpose_scale<false><<<dim3((unsigned)rb, (unsigned)cbk), block, 0, st>>>(in, R, C, scale, out);
    } else {
        if (rb > 65535) throw Error(5, "matrix too large for the transpose kernel");
        k_transpose_scale<true><<<dim3((unsigned)cbk, (unsigned)rb), block, 0, st>>>(in, R, C, scale, out);
    }
    CDL_CUDA(cudaGetLastError());
}

void launch_scale(double* x, int64_t n, double s, cudaStream_t st) {
    if (n <= 0) return;
    k_scale<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, s);
    CDL_CUDA(cudaGetLastError());
}

void launch_geweke(const double* prob_all, int64_t cols, int64_t n_rows, unsigned long long* counts2,
                   cudaStream_t st) {
    CDL_CUDA(cudaMemsetAsync(counts2, 0, 2 * sizeof(unsigned long long), st));
    if (cols <= TILED_MAX_COLS)
        k_geweke_tiled<<<dim3((unsigned)((cols + GW_TILE - 1) / GW_TILE), 1), dim3(GW_TILE, GW_TILE), 0, st>>>(
            prob_all, nullptr, cols, n_rows, counts2);
    else
        k_geweke<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(prob_all, cols, n_rows, counts2);
    CDL_CUDA(cudaGetLastError());
}

// The test on `n_traces` traces of the same shape in one launch (the chains of a small donor at a convergence check):
// traces_dev is a DEVICE array of their pointers, counts gets 2 * n_traces values (tested columns, columns with |Z| <= 2).
// Returns false (nothing launched) when the traces are too wide for the tiled kernel: the caller tests them one by one.
bool launch_geweke_batch(const double* const* traces_dev, int n_traces, int64_t cols, int64_t n_rows,
                         unsigned long long* counts, cudaStream_t st) {
    if (cols > TILED_MAX_COLS || n_traces > 65535) return false;
    if (n_traces <= 0) return true;
    CDL_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long) * (size_t)n_traces, st));
    k_geweke_tiled<<<dim3((unsigned)((cols + GW_TILE - 1) / GW_TILE), (unsigned)n_traces), dim3(GW_TILE, GW_TILE), 0, st>>>(
        nullptr, traces_dev, cols, n_rows, counts);
    CDL_CUDA(cudaGetLastError());
    return true;
}

void launch_geweke_sums(const void* s1a, const void* s2a, const void* s1b, const void* s2b, bool fp32, const double* s1n,
                        const double* s2n, int64_t cols, int64_t n_rows, unsigned long long* counts2, cudaStream_t st) {
    CDL_CUDA(cudaMemsetAsync(counts2, 0, 2 * sizeof(unsigned long long), st));
    const unsigned grid = (unsigned)((cols + 255) / 256);
    if (fp32)
        k_geweke_sums<float><<<grid, 256, 0, st>>>((const float*)s1a, (const float*)s2a, (const float*)s1b,
                                                   (const float*)s2b, s1n, s2n, cols, n_rows, counts2);
    else
        k_geweke_sums<double><<<grid, 256, 0, st>>>((const double*)s1a, (const double*)s2a, (const double*)s1b,
                                                    (const double*)s2b, s1n, s2n, cols, n_rows, counts2);
    CDL_CUDA(cudaGetLastError());
}

// ---- segment sums of the streaming stop rule (api.cu): store + clear, fold, add up ----
// fp32 segments keep the sum of SQUARES as its square root: a clone column that sits at 1e-25 has squares of 1e-50,
// far below fp32's range (1e-38, denormals to 1e-45) but well inside the 1e-50 floor of the Z statistic's
// denominator, where the verdict is decided; the root is representable down to sums of 1e-76.
namespace {
template <typename S> struct SegCodec {
    static __device__ __forceinline__ S put(double v, bool sq) { return (S)(sq ? sqrt(v) : v); }
    static __device__ __forceinline__ double get(S v, bool sq) { const double d = (double)v; return sq ? d * d : d; }
};
template <> struct SegCodec<double> {
    static __device__ __forceinline__ double put(double v, bool) { return v; }
    static __device__ __forceinline__ double get(double v, bool) { return v; }
};
template <typename S>
__global__ void k_segment_store(double* __restrict__ acc1, double* __restrict__ acc2, S* __restrict__ out1,
                                S* __restrict__ out2, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (out1) out1[i] = SegCodec<S>::put(acc1[i], false);
    if (out2) out2[i] = SegCodec<S>::put(acc2[i], true);
    acc1[i] = 0.0;
    acc2[i] = 0.0;
}
template <typename S>
__global__ void k_segment_add(double* __restrict__ dst, const S* __restrict__ src, bool squares, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] += SegCodec<S>::get(src[i], squares);
}
// out = extra + sum of `count` segment arrays, in list order (fixed rounding)
template <typename S>
__global__ void k_segment_sum(const void* const* __restrict__ ptrs, int count, bool squares,
                              const double* __restrict__ extra, int64_t n, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int q = 0; q < count; ++q) s += SegCodec<S>::get(reinterpret_cast<const S*>(ptrs[q])[i], squares);
    out[i] = s + (extra ? extra[i] : 0.0);
}
}  // namespace

void launch_segment_store(double* acc1, double* acc2, void* out1, void* out2, bool fp32, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (fp32) k_segment_store<float><<<grid, 256, 0, st>>>(acc1, acc2, (float*)out1, (float*)out2, n);
    else k_segment_store<double><<<grid, 256, 0, st>>>(acc1, acc2, (double*)out1, (double*)out2, n);
    CDL_CUDA(cudaGetLastError());
}

void launch_segment_add(double* dst, const void* src, bool fp32, bool squares, int64_t n, cudaStream_t st) {
    if (n <= 0 || !src) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (fp32) k_segment_add<float><<<grid, 256, 0, st>>>(dst, (const float*)src, squares, n);
    else k_segment_add<double><<<grid, 256, 0, st>>>(dst, (const double*)src, squares, n);
    CDL_CUDA(cudaGetLastError());
}

void launch_segment_sum(const void* const* ptrs, int count, bool fp32, bool squares, const double* extra, int64_t n,
                        double* out, cudaStream_t st) {
    if (n <= 0) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (fp32) k_segment_sum<float><<<grid, 256, 0, st>>>(ptrs, count, squares, extra, n, out);
    else k_segment_sum<double><<<grid, 256, 0, st>>>(ptrs, count, squares, extra, n, out);
    CDL_CUDA(cudaGetLastError());
}

void launch_window_mean(const double* s1n, const void* s1u, bool fp32, int64_t n, double inv, double* out, cudaStream_t st) {
    if (n <= 0) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (fp32) k_window_mean<float><<<grid, 256, 0, st>>>(s1n, (const float*)s1u, n, inv, out);
    else k_window_mean<double><<<grid, 256, 0, st>>>(s1n, (const double*)s1u, n, inv, out);
    CDL_CUDA(cudaGetLastError());
}

template <typename T>
static void launch_col_means_t(const T* X, int64_t cols, int64_t r0, int64_t r1, double* out, cudaStream_t st) {
    if (cols <= TILED_MAX_COLS)
        k_col_means_tiled<T><<<(unsigned)((cols + GW_TILE - 1) / GW_TILE), dim3(GW_TILE, GW_TILE), 0, st>>>(X, cols, r0, r1, out);
    else
        k_col_means<T><<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(X, cols, r0, r1, out);
    CDL_CUDA(cudaGetLastError());
}
void launch_col_means_d(const double* X, int64_t cols, int64_t r0, int64_t r1, double* out, cudaStream_t st) {
    launch_col_means_t<double>(X, cols, r0, r1, out, st);
}
void launch_col_means_u8(const uint8_t* X, int64_t cols, int64_t r0, int64_t r1, double* out, cudaStream_t st) {
    launch_col_means_t<uint8_t>(X, cols, r0, r1, out, st);
}

void launch_prob_variant(const Donor& d, int wise, const double* theta0_all, const double* theta1_all,
                         int64_t n_el, int64_t r0, int64_t r1, double* out_dense, cudaStream_t st) {
    const int64_t n = d.N * d.M;
    // uncovered entries: dbinom(0, 0, .) = 1 on both sides -> 0.5 for global / variant; NA for element, where the
    // reference only assigns the covered entries (R/clone_id.R:617-621)
    k_fill<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out_dense, n, wise == 2 ? (double)NAN : 0.5);
    const unsigned grid = warp_grid(d.M, 256);
    if (d.vidx32)
        k_prob_variant<uint32_t><<<grid, 256, 0, st>>>(d.c_rowgrp, (const uint32_t*)d.c_vidx, d.c_a, d.c_b, d.M, d.N,
                                                       wise, theta0_all, theta1_all, n_el, r0, r1, d.c_elem0, out_dense);
    else
        k_prob_variant<uint16_t><<<grid, 256, 0, st>>>(d.c_rowgrp, (const uint16_t*)d.c_vidx, d.c_a, d.c_b, d.M, d.N,
                                                       wise, theta0_all, theta1_all, n_el, r0, r1, d.c_elem0, out_dense);
    CDL_CUDA(cudaGetLastError());
}

double loglik_general(const Donor& d, const double* Config_dev, int K, const int32_t* assign_dev,
                      const double* prob_dev, double theta0, const double* theta1_dev, int64_t n_theta1,
                      cudaStream_t st) {
    double* cell = nullptr;
    double* out = nullptr;
    CDL_CUDA(cudaMalloc(&cell, sizeof(double) * (size_t)d.M));
    CDL_CUDA(cudaMalloc(&out, sizeof(double)));
    const unsigned grid = warp_grid(d.M, 256);
    const double l0 = log(theta0), l0c = log(1.0 - theta0);
    if (d.vidx32)
        k_loglik<uint32_t><<<grid, 256, 0, st>>>(d.c_rowgrp, (const uint32_t*)d.c_vidx, d.c_a, d.c_b, d.M, d.N, K,
                                                 Config_dev, assign_dev, prob_dev, l0, l0c, theta1_dev, n_theta1,
                                                 (n_theta1 != 1 && n_theta1 != d.N) ? d.c_elem0 : nullptr, cell);
    else
        k_loglik<uint16_t><<<grid, 256, 0, st>>>(d.c_rowgrp, (const uint16_t*)d.c_vidx, d.c_a, d.c_b, d.M, d.N, K,
                                                 Config_dev, assign_dev, prob_dev, l0, l0c, theta1_dev, n_theta1,
                                                 (n_theta1 != 1 && n_theta1 != d.N) ? d.c_elem0 : nullptr, cell);
    k_sum_fixed<<<1, 256, 0, st>>>(cell, d.M, out);
    double h = 0.0;
    cudaError_t e = cudaMemcpyAsync(&h, out, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(cell);
    cudaFree(out);
    CDL_CUDA(e);
    CDL_CUDA(cudaGetLastError());
    return h + d.W_log;
}

void em_stats(const Donor& d, const uint32_t* cfg_mask, int K, uint32_t* S3, uint32_t* S4, uint32_t* TA,
              uint32_t* TB, int sm_count, cudaStream_t st) {
    (void)sm_count;
    const unsigned grid = warp_grid(d.M, 256);
    if (d.vidx32)
        k_em_stats<uint32_t><<<grid, 256, 0, st>>>(d.c_rowgrp, (const uint32_t*)d.c_vidx, d.c_a, d.c_b, d.M, K,
                                                   cfg_mask, S3, S4, TA, TB);
    else
        k_em_stats<uint16_t><<<grid, 256, 0, st>>>(d.c_rowgrp, (const uint16_t*)d.c_vidx, d.c_a, d.c_b, d.M, K,
                                                   cfg_mask, S3, S4, TA, TB);
    CDL_CUDA(cudaGetLastError());
}

void em_iteration(int64_t M, int K, const uint32_t* S3, const uint32_t* S4, const uint32_t* TA,
                  const uint32_t* TB, const double* logPsi_dev, double theta0, double theta1, double* prob_out,
                  double* sums5_dev, double* partials, cudaStream_t st) {
    const int64_t nblocks = (M + 255) / 256;
    k_em_iter<<<(unsigned)nblocks, 256, 0, st>>>(M, K, S3, S4, TA, TB, logPsi_dev, log(theta0), log(1.0 - theta0),
                                                 log(theta1), log(1.0 - theta1), prob_out, partials);
    k_em_final<<<1, 256, 0, st>>>(partials, nblocks, sums5_dev);
    CDL_CUDA(cudaGetLastError());
}

}  // namespace cdl
