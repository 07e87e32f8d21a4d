// Small donors (the reference's own scale: example_donor is 34 variants x 428 cells, nnz 1874): the WHOLE sweep of
// a chain -- logLik, softmax, draw, read sums, Config flips, theta draws, logLik trace (R/clone_id.R:466-577) -- in
// ONE CTA, many sweeps per launch, one CTA per chain.
//
// At this size the three-kernel sweep of gibbs.cu is pure launch latency (~5 launches of a few microseconds of
// work each, ~30-40 us per sweep, and independent chains queue behind one another).  Here the per-variant state
// (ln theta1, Config masks, the N x K read-sum tables) lives in shared memory for the whole launch, the phases
// are separated by __syncthreads instead of kernel boundaries, and the chains of a run are the blocks of one
// grid.  Philox keys, Beta sampler and arithmetic are those of k_cells / k_stats / k_table, so the chain is the
// one the three-kernel path produces (up to the summation order inside a logit).  The Geweke stop rule stays on
// the host side of the launch boundary: the driver launches `check_every` sweeps at a time (api.cu).
#include "internal.h"
#include "cells.cuh"

namespace cdl {

namespace {

constexpr int SMALL_THREADS = 512;      // up to 32 fp64 logits per lane: 128 registers per thread

struct SmallParams {
    const uint32_t* rowgrp; const void* vidx; const uint16_t* ca; const uint16_t* cb;
    int vidx32;
    int64_t M, N, cell_offset, n_el, T, n_buin_planned;
    int K, wise, relax, keep_base, relax_fixed_set, min_iter, exact, chain0;
    double prior0[2], prior1[2], relax_prior[2];
    const double* prior1_mat;
    double logPsi[KMAX];
    const uint32_t* cfg0;
    double W_log;
    PhiloxKey key;
    uint32_t it_begin, it_end;          // sweeps it_begin .. it_end (1-based trace rows) in this launch
    const ChainDev* chains;             // [gridDim.x]
    const int* active;                  // [gridDim.x] 0: the chain has stopped, skip it
};

__device__ __forceinline__ double block_sum_d_small(double v, double* sm) {
    v = warp_sum_d(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = 0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) r += sm[q];     // fixed order: deterministic
    return r;
}

template <int KP>
__global__ void __launch_bounds__(SMALL_THREADS, 1) k_small_chain(const SmallParams P) {
    if (!P.active[blockIdx.x]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ChainDev c = P.chains[blockIdx.x];
    const uint32_t chain = (uint32_t)(P.chain0 + (int)blockIdx.x);
    const int N = (int)P.N, K = P.K, NK = N * K;
    const int64_t M = P.M;
    const int n_tab = P.wise == 1 ? N : 1;
    // ---- shared-memory state of the chain ----
    double* s_l1 = reinterpret_cast<double*>(smem_raw);            // [n_tab] ln theta1
    double* s_l1c = s_l1 + n_tab;                                  // [n_tab] ln(1 - theta1)
    double* s_gam = s_l1c + n_tab;                                 // [2 * n_tab + 6] Gamma draws of the sweep
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_gam + 2 * n_tab + 6);   // [N] Config_new rows
    uint32_t* s_new = s_mask + N;                                  // [N] rows being drawn
    uint32_t* s_cfg0 = s_new + N;                                  // [N] input Config rows
    uint32_t* s_s3 = s_cfg0 + N;                                   // [N] sum_k SA C'
    uint32_t* s_s4 = s_s3 + N;                                     // [N]
    uint32_t* s_sa = s_s4 + N;                                     // [N*K] SA
    uint32_t* s_sb = s_sa + NK;                                    // [N*K] SB
    __shared__ double s_red[SMALL_THREADS / 32];
    __shared__ double s_l0, s_l0c, s_relax, s_ll;
    __shared__ unsigned long long s_tot[5];                        // S1, S2, S3 total, S4 total, diff1
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    for (int i = tid; i < n_tab; i += blockDim.x) { s_l1[i] = c.l1[i]; s_l1c[i] = c.l1c[i]; }
    for (int i = tid; i < N; i += blockDim.x) { s_mask[i] = c.mask[i]; s_cfg0[i] = P.cfg0[i]; }
    if (tid == 0) { s_l0 = c.scal[SC_L0]; s_l0c = c.scal[SC_L0C]; s_relax = c.scal[SC_RELAX]; }
    __syncthreads();

    for (uint32_t it = P.it_begin; it <= P.it_end; ++it) {
        const int64_t row = (int64_t)it - 1;
        for (int t = tid; t < NK; t += blockDim.x) { s_sa[t] = 0u; s_sb[t] = 0u; }
        for (int i = tid; i < N; i += blockDim.x) { s_s3[i] = 0u; s_s4[i] = 0u; s_new[i] = 0u; }
        if (tid < 5) s_tot[tid] = 0ull;
        __syncthreads();
        const double l0 = s_l0, l0c = s_l0c;
        // ---- cells: logLik -> softmax -> draw (:470-497), one lane per cell; read sums of the drawn label ----
        double* prob_row = c.prob_all ? c.prob_all + row * M * K : nullptr;
        double* prob_acc = (!P.exact && (int64_t)it >= P.n_buin_planned) ? c.prob_acc : nullptr;
        uint8_t* assign_row = c.assign_all ? c.assign_all + row * M : nullptr;
        for (int64_t j0 = (int64_t)warp * 32; j0 < M; j0 += (int64_t)nwarps * 32) {
            const int64_t j = j0 + lane;
            const bool valid = j < M;
            int64_t e0 = 0, e1 = 0;
            if (valid) { e0 = (int64_t)P.rowgrp[j] * GROUP; e1 = (int64_t)P.rowgrp[j + 1] * GROUP; }
            double acc[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = 0.0;
            for (int64_t e = e0; e < e1; ++e) {
                const uint32_t a = P.ca[e], b = P.cb[e];
                if ((a | b) == 0) continue;
                const uint32_t v = P.vidx32 ? reinterpret_cast<const uint32_t*>(P.vidx)[e]
                                            : reinterpret_cast<const uint16_t*>(P.vidx)[e];
                const int t = P.wise == 1 ? (int)v : 0;
                const double del = fma((double)a, s_l1[t] - l0, (double)b * (s_l1c[t] - l0c));
                const uint32_t m = s_mask[v];
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[k] += ((m >> k) & 1u) ? del : 0.0;
            }
            lane_softmax<KP>(acc, K, P.logPsi);
            if (valid) {
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    if (k < K) {
                        if (prob_row) prob_row[j * K + k] = acc[k];
                        if (prob_acc) prob_acc[j * K + k] += acc[k];
                    }
                }
            }
            const double u = valid ? philox_uniform(P.key, chain, it, STREAM_ASSIGN, (uint32_t)(P.cell_offset + j)) : 0.5;
            const int sel = lane_draw<KP>(acc, K, u, valid);
            if (valid) {
                c.z[j] = (uint8_t)sel;
                if (assign_row) assign_row[j] = (uint8_t)(sel + 1);
                for (int64_t e = e0; e < e1; ++e) {
                    const uint32_t a = P.ca[e], b = P.cb[e];
                    if ((a | b) == 0) continue;
                    const uint32_t v = P.vidx32 ? reinterpret_cast<const uint32_t*>(P.vidx)[e]
                                                : reinterpret_cast<const uint16_t*>(P.vidx)[e];
                    if (a) atomicAdd(&s_sa[v * K + sel], a);
                    if (b) atomicAdd(&s_sb[v * K + sel], b);
                }
            }
        }
        __syncthreads();
        // ---- table: Config flips (:519-535), theta statistics (:546-562), logLik trace with the OLD theta ----
        const double r = s_relax;
        const double odd1 = log(1.0 - r) - log(r), odd0 = log(r) - log(1.0 - r);
        double ll = 0.0;
        unsigned long long S1 = 0, S2 = 0, S3t = 0, S4t = 0, d1 = 0;
        for (int t = tid; t < NK; t += blockDim.x) {
            const int i = t / K, k = t - i * K;
            const int tt = P.wise == 1 ? i : 0;
            const double l1 = s_l1[tt], l1c = s_l1c[tt];
            const uint32_t sa = s_sa[t], sb = s_sb[t];
            const double SA = (double)sa, SB = (double)sb;
            const uint32_t cbit = (s_cfg0[i] >> k) & 1u;
            uint32_t bit = cbit;
            if (P.relax) {
                double prior;
                if (P.keep_base && k == 0) prior = cbit ? INFINITY : -INFINITY;
                else prior = cbit ? odd1 : odd0;
                double x = SA * (l1 - l0) + SB * (l1c - l0c) + prior;
                x = x > 50.0 ? 50.0 : (x < -50.0 ? -50.0 : x);
                const double ex = exp(x);
                const double p = ex / (ex + 1.0);
                const uint32_t idx = (uint32_t)(i + (int64_t)k * N);
                bit = (uint32_t)rbinom1(p, philox_uniform(P.key, chain, it, STREAM_CONFIG, idx));
            }
            c.config_all[row * NK + i + (int64_t)k * N] = (uint8_t)bit;
            if (bit) {
                atomicOr(&s_new[i], 1u << k);
                if (sa) atomicAdd(&s_s3[i], sa);
                if (sb) atomicAdd(&s_s4[i], sb);
                S3t += sa; S4t += sb;
                ll += SA * l1 + SB * l1c;
            } else {
                S1 += sa; S2 += sb;
                ll += SA * l0 + SB * l0c;
            }
            if (k >= 1) d1 += (bit != cbit);
        }
        S1 = warp_sum_u64(S1); S2 = warp_sum_u64(S2); S3t = warp_sum_u64(S3t); S4t = warp_sum_u64(S4t);
        d1 = warp_sum_u64(d1);
        if (lane == 0) {
            atomicAdd(&s_tot[0], S1); atomicAdd(&s_tot[1], S2); atomicAdd(&s_tot[2], S3t); atomicAdd(&s_tot[3], S4t);
            atomicAdd(&s_tot[4], d1);
        }
        const double tl = block_sum_d_small(ll, s_red);        // (synchronises: the tables above are complete)
        if (tid == 0) s_ll = tl;
        __syncthreads();
        // ---- every Beta draw of the sweep = two Gamma jobs (same Philox sub-streams as k_table) ----
        const bool learn_next = P.relax && !P.relax_fixed_set && ((double)(it + 1) > 0.1 * P.min_iter + 5.0) &&
                                (int64_t)it < P.T;
        const int n_var_jobs = P.wise == 1 ? 2 * N : 0;
        for (int job = tid; job < n_var_jobs + 6; job += blockDim.x) {
            double shape = -1.0;
            uint32_t stream = 0, idx = 0, pit = it;
            if (job < n_var_jobs) {
                const int iv = job >> 1;
                const double pa = P.prior1_mat ? P.prior1_mat[iv] : P.prior1[0];
                const double pb = P.prior1_mat ? P.prior1_mat[P.n_el + iv] : P.prior1[1];
                shape = (job & 1) ? pb + (double)s_s4[iv] : pa + (double)s_s3[iv];
                stream = STREAM_THETA1; idx = (uint32_t)iv;
            } else {
                const int q = job - n_var_jobs;               // 0,1 theta0; 2,3 next relax rate; 4,5 global theta1
                if (q < 2) { shape = P.prior0[q] + (double)s_tot[q]; stream = STREAM_THETA0; }
                else if (q < 4) {
                    if (learn_next) {
                        const double dd1 = (double)s_tot[4];
                        shape = (q == 2) ? P.relax_prior[0] + dd1 : P.relax_prior[1] + (double)N * (double)(K - 1) - dd1;
                        stream = STREAM_RELAX; pit = it + 1;
                    }
                } else if (P.wise == 0) {
                    const double pa = P.prior1_mat ? P.prior1_mat[0] : P.prior1[0];
                    const double pb = P.prior1_mat ? P.prior1_mat[1] : P.prior1[1];
                    shape = (q == 4) ? pa + (double)s_tot[2] : pb + (double)s_tot[3];
                    stream = STREAM_THETA1;
                }
            }
            if (shape >= 0.0) {
                PhiloxSeq sq(P.key, chain, pit, stream, idx, (uint32_t)(job & 1) << 31);
                s_gam[job] = gamma_mt(shape, sq);
            }
        }
        __syncthreads();
        if (P.wise == 1) {
            for (int i = tid; i < N; i += blockDim.x) {
                const double th = beta_from_gammas(s_gam[2 * i], s_gam[2 * i + 1]);
                c.theta1_all[row * P.n_el + i] = th;
                s_l1[i] = log(th);
                s_l1c[i] = log(1.0 - th);
            }
        }
        for (int i = tid; i < N; i += blockDim.x) s_mask[i] = s_new[i];
        if (tid == 0) {
            c.loglik_all[row] = s_ll + P.W_log;
            const double th0 = beta_from_gammas(s_gam[n_var_jobs], s_gam[n_var_jobs + 1]);
            c.theta0_all[row] = th0;
            s_l0 = log(th0);
            s_l0c = log(1.0 - th0);
            c.scal[SC_THETA0] = th0;
            if (P.wise == 0) {
                const double th = beta_from_gammas(s_gam[n_var_jobs + 4], s_gam[n_var_jobs + 5]);
                c.theta1_all[row] = th;
                s_l1[0] = log(th);
                s_l1c[0] = log(1.0 - th);
            }
            if (learn_next) {
                const double rr = beta_from_gammas(s_gam[n_var_jobs + 2], s_gam[n_var_jobs + 3]);
                s_relax = rr;
                c.relax_all[row + 1] = rr;
            }
        }
        __syncthreads();
    }
    // ---- state back to global memory (the next launch, or the three-kernel path / bench hook, resumes from it) ----
    for (int i = tid; i < n_tab; i += blockDim.x) { c.l1[i] = s_l1[i]; c.l1c[i] = s_l1c[i]; }
    for (int i = tid; i < N; i += blockDim.x) c.mask[i] = s_mask[i];
    if (tid == 0) {
        c.scal[SC_L0] = s_l0; c.scal[SC_L0C] = s_l0c; c.scal[SC_RELAX] = s_relax;
        c.scal[SC_ODD1] = log(1.0 - s_relax) - log(s_relax);
    }
}

}  // namespace

size_t small_smem_bytes(int64_t N, int K, int wise) {
    const size_t n_tab = wise == 1 ? (size_t)N : 1;
    return sizeof(double) * (2 * n_tab + 2 * n_tab + 6) + sizeof(uint32_t) * (5 * (size_t)N + 2 * (size_t)N * K) + 64;
}

bool small_path_fits(const Donor& d, int K, int wise) {
    return d.M <= 65536 && d.nnz <= (1 << 18) && d.N * (int64_t)K <= 16384 && small_smem_bytes(d.N, K, wise) <= 200 * 1024 &&
           d.max_variant_depth < (1ull << 32);
}

void launch_small_chains(const Donor& d, const SweepArgs& a, const cdl_small_run& r, cudaStream_t st) {
    SmallParams P{};
    P.rowgrp = d.c_rowgrp; P.vidx = d.c_vidx; P.ca = d.c_a; P.cb = d.c_b; P.vidx32 = d.vidx32 ? 1 : 0;
    P.M = d.M; P.N = d.N; P.cell_offset = d.cell_offset; P.n_el = a.wise == 1 ? d.N : 1; P.T = a.T;
    P.n_buin_planned = r.n_buin_planned;
    P.K = a.K; P.wise = a.wise; P.relax = a.relax; P.keep_base = a.keep_base; P.relax_fixed_set = a.relax_fixed_set;
    P.min_iter = r.min_iter; P.exact = r.exact; P.chain0 = r.chain0;
    for (int q = 0; q < 2; ++q) { P.prior0[q] = a.prior0[q]; P.prior1[q] = a.prior1[q]; P.relax_prior[q] = a.relax_prior[q]; }
    P.prior1_mat = a.prior1_mat;
    for (int k = 0; k < KMAX; ++k) P.logPsi[k] = a.logPsi[k];
    P.cfg0 = a.cfg0; P.W_log = a.W_log; P.key = a.key;
    P.it_begin = r.it_begin; P.it_end = r.it_end; P.chains = r.chains_dev; P.active = r.active_dev;
    const size_t smem = small_smem_bytes(d.N, a.K, a.wise);
    const int KP = a.K <= 4 ? 4 : (a.K <= 8 ? 8 : (a.K <= 16 ? 16 : 32));
#define CDL_SMALL_LAUNCH(KPV)                                                                                   \
    do {                                                                                                        \
        auto kern = k_small_chain<KPV>;                                                                         \
        CDL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        kern<<<(unsigned)r.n_chain, SMALL_THREADS, smem, st>>>(P);                                              \
    } while (0)
    switch (KP) {
        case 4: CDL_SMALL_LAUNCH(4); break;
        case 8: CDL_SMALL_LAUNCH(8); break;
        case 16: CDL_SMALL_LAUNCH(16); break;
        default: CDL_SMALL_LAUNCH(32); break;
    }
#undef CDL_SMALL_LAUNCH
    CDL_CUDA(cudaGetLastError());
}

}  // namespace cdl
