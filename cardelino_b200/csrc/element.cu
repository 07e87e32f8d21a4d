// wise = "element" (R/clone_id.R:413-416): one theta1 per COVERED ENTRY.
//
// The reference only defines this mode on fully covered data -- at an uncovered entry theta1 is NA and the NA
// goes through `%*%` into Config_new; here uncovered entries are skipped, which is the
// same thing wherever the reference is defined.  Per-entry parameters mean per-entry state (log theta1 and
// log(1 - theta1) beside every count of the row layout), a per-entry Beta draw every sweep and an
// iterations x nnz trace, so this path is meant for the small donors the mode is usable on; it is written for
// clarity (one warp per cell on the row layout, fp64 atomics for the per-(variant, clone) log-odds table), not
// tuned like the global / variant path.  On cell shards every rank owns the parameters of its own entries (prior1
// matrix rows and theta1 trace columns are the rank's entries); the two N x K tables are all-reduced with NCCL
// (api.cu::one_sweep) and the Philox index of an entry is its position in the whole donor.
//
//   k_el_cells   logLik with per-entry weights -> softmax -> categorical draw            (:470-497)
//   k_el_stats   SA / SB (integer) and SD[i,k] = sum_{j: z_j = k} A ln th1_ij + B ln(1 - th1_ij)   (:525-528)
//   k_table      (gibbs.cu, element branch) Config flips from SD, theta0 / relax draws, logLik trace
//   k_el_theta   theta1_ij ~ Beta(prior + A C', prior + B C') for every covered entry     (:557-571)
#include "internal.h"
#include "cells.cuh"

namespace cdl {

namespace {

constexpr int EL_THREADS = 256;

inline unsigned el_grid(int64_t M) {
    const int64_t blocks = (M * 32 + EL_THREADS - 1) / EL_THREADS;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(blocks, 148 * 8));
}

struct ElParams {
    const uint32_t* rowgrp;
    const void* vidx;
    const uint16_t* ca;
    const uint16_t* cb;
    const int64_t* elem0;      // [M+1] element index of each cell's first covered entry
    int64_t M, N, cell_offset, n_el;
    int64_t el_offset;         // covered entries of the ranks before this one (cell shards)
    int K, vidx32;
    double* el_l1;             // [padded entries] ln theta1
    double* el_l1c;            //                  ln(1 - theta1)
    const double* scal;
    const uint32_t* mask;
    double logPsi[KMAX];
    double* prob_row; double* prob_acc; uint8_t* z; uint8_t* assign_row;
    unsigned long long* sab; double* sd1;
    double* theta1_row;        // [n_el] trace row to fill
    double prior1[2]; const double* prior1_mat;
    const double* rp_u; const double* rp_theta1;
    PhiloxKey key; uint32_t chain, it;
};

__device__ __forceinline__ uint32_t el_variant(const ElParams& P, int64_t e) {
    return P.vidx32 ? reinterpret_cast<const uint32_t*>(P.vidx)[e] : reinterpret_cast<const uint16_t*>(P.vidx)[e];
}

template <int KP>
__global__ void __launch_bounds__(EL_THREADS) k_el_cells(const ElParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double l0 = P.scal[SC_L0], l0c = P.scal[SC_L0C];
    for (int64_t j = warp; j < P.M; j += nwarps) {
        const int64_t e0 = (int64_t)P.rowgrp[j] * GROUP, e1 = (int64_t)P.rowgrp[j + 1] * GROUP;
        double acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.0;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const uint32_t a = P.ca[e], b = P.cb[e];
            if ((a | b) == 0) continue;
            const uint32_t m = P.mask[el_variant(P, e)];
            const double del = (double)a * (P.el_l1[e] - l0) + (double)b * (P.el_l1c[e] - l0c);
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] += ((m >> k) & 1u) ? del : 0.0;
        }
#pragma unroll
        for (int k = 0; k < KP; ++k)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[k] += shfl_xor_d(acc[k], off);   // every lane: full sums
        lane_softmax<KP>(acc, P.K, P.logPsi);
        if (lane == 0) {
            for (int k = 0; k < P.K; ++k) {
                if (P.prob_row) P.prob_row[j * P.K + k] = acc[k];
                if (P.prob_acc) P.prob_acc[j * P.K + k] += acc[k];
            }
        }
        const double u = P.rp_u ? P.rp_u[j] : philox_uniform(P.key, P.chain, P.it, STREAM_ASSIGN,
                                                             (uint32_t)(P.cell_offset + j));
        const int sel = lane_draw<KP>(acc, P.K, u, true);
        if (lane == 0) {
            P.z[j] = (uint8_t)sel;
            if (P.assign_row) P.assign_row[j] = (uint8_t)(sel + 1);
        }
    }
}

__global__ void __launch_bounds__(EL_THREADS) k_el_stats(const ElParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < P.M; j += nwarps) {
        const int64_t e0 = (int64_t)P.rowgrp[j] * GROUP, e1 = (int64_t)P.rowgrp[j + 1] * GROUP;
        const int zj = P.z[j];
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const unsigned long long a = P.ca[e], b = P.cb[e];
            if ((a | b) == 0) continue;
            const int64_t t = (int64_t)el_variant(P, e) * P.K + zj;
            atomicAdd(&P.sab[t], a | (b << 32));
            atomicAdd(&P.sd1[t], (double)a * P.el_l1[e] + (double)b * P.el_l1c[e]);
        }
    }
}

// theta1 of every covered entry: initial draw (INIT, :462) or the sweep's posterior draw (:557-571, C' = new Config)
template <bool INIT>
__global__ void __launch_bounds__(EL_THREADS) k_el_theta(const ElParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < P.M; j += nwarps) {
        const int64_t e0 = (int64_t)P.rowgrp[j] * GROUP, e1 = (int64_t)P.rowgrp[j + 1] * GROUP;
        const int zj = INIT ? 0 : P.z[j];
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const uint32_t a = P.ca[e], b = P.cb[e];
            if ((a | b) == 0) { P.el_l1[e] = 0.0; P.el_l1c[e] = 0.0; continue; }
            const int64_t el = P.elem0[j] + (e - e0);
            double th;
            if (P.rp_theta1) th = clamp_theta(P.rp_theta1[el]);
            else {
                double pa = P.prior1_mat ? P.prior1_mat[el] : P.prior1[0];
                double pb = P.prior1_mat ? P.prior1_mat[P.n_el + el] : P.prior1[1];
                if (!INIT && ((P.mask[el_variant(P, e)] >> zj) & 1u)) { pa += (double)a; pb += (double)b; }
                PhiloxSeq s(P.key, P.chain, P.it, STREAM_THETA1, (uint32_t)(P.el_offset + el));
                th = beta_draw(pa, pb, s);
            }
            P.theta1_row[el] = th;
            P.el_l1[e] = log(th);
            P.el_l1c[e] = log(1.0 - th);
        }
    }
}

ElParams el_params(const Donor& d, const SweepArgs& a, ChainDev& c) {
    ElParams P{};
    P.rowgrp = d.c_rowgrp; P.vidx = d.c_vidx; P.ca = d.c_a; P.cb = d.c_b; P.elem0 = d.c_elem0;
    P.M = d.M; P.N = d.N; P.cell_offset = d.cell_offset; P.n_el = d.nnz; P.el_offset = a.el_offset; P.K = a.K; P.vidx32 = d.vidx32 ? 1 : 0;
    P.el_l1 = c.el_l1; P.el_l1c = c.el_l1c; P.scal = c.scal; P.mask = c.mask;
    for (int k = 0; k < KMAX; ++k) P.logPsi[k] = a.logPsi[k];
    const int64_t row = (int64_t)a.it - 1;
    P.prob_row = c.prob_all ? c.prob_all + row * d.M * a.K : nullptr;
    P.prob_acc = c.prob_acc;
    P.z = c.z;
    P.assign_row = c.assign_all ? c.assign_all + row * d.M : nullptr;
    P.sab = c.sab; P.sd1 = c.sd1;
    P.theta1_row = c.theta1_all + row * d.nnz;
    P.prior1[0] = a.prior1[0]; P.prior1[1] = a.prior1[1]; P.prior1_mat = a.prior1_mat;
    P.rp_u = a.rp_u_assign; P.rp_theta1 = a.rp_theta1;
    P.key = a.key; P.chain = a.chain; P.it = a.it;
    return P;
}

}  // namespace

void launch_el_init(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st) {
    const ElParams P = el_params(d, a, c);
    k_el_theta<true><<<el_grid(d.M), EL_THREADS, 0, st>>>(P);
    CDL_CUDA(cudaGetLastError());
}

void launch_el_cells(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st) {
    const ElParams P = el_params(d, a, c);
    const unsigned grid = el_grid(d.M);
    const int KP = a.K <= 4 ? 4 : (a.K <= 8 ? 8 : (a.K <= 16 ? 16 : 32));
    switch (KP) {
        case 4: k_el_cells<4><<<grid, EL_THREADS, 0, st>>>(P); break;
        case 8: k_el_cells<8><<<grid, EL_THREADS, 0, st>>>(P); break;
        case 16: k_el_cells<16><<<grid, EL_THREADS, 0, st>>>(P); break;
        default: k_el_cells<32><<<grid, EL_THREADS, 0, st>>>(P); break;
    }
    CDL_CUDA(cudaGetLastError());
}

void launch_el_stats(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st) {
    CDL_CUDA(cudaMemsetAsync(c.sab, 0, sizeof(unsigned long long) * (size_t)(d.N * a.K), st));
    CDL_CUDA(cudaMemsetAsync(c.sd1, 0, sizeof(double) * (size_t)(d.N * a.K), st));
    const ElParams P = el_params(d, a, c);
    k_el_stats<<<el_grid(d.M), EL_THREADS, 0, st>>>(P);
    CDL_CUDA(cudaGetLastError());
}

void launch_el_theta(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st) {
    const ElParams P = el_params(d, a, c);
    k_el_theta<false><<<el_grid(d.M), EL_THREADS, 0, st>>>(P);
    CDL_CUDA(cudaGetLastError());
}

}  // namespace cdl
