// The Gibbs sweep of clone_id_Gibbs (R/clone_id.R:466-593) as three kernels on sm_100a:
//
//   k_cells  one pass over the cell-major layout: logLik thin SpMM (:470-478) -> softmax (:479-481)
//            -> categorical draw with sample()'s sorted-CDF walk (:492-497).
//   k_stats  one pass over the variant-major layout: SA[i,k] = sum_{j: z_j = k} A_ij, SB likewise.
//            These N x K integer tables carry everything the rest of the sweep needs:
//            they replace the 4K dense S-lists (:394-403,537-542), the two N x M x K products of the
//            relax step (:525-528) and the weighted sums of the theta step (:546-562).
//   k_table  N x K work: relax-rate Beta draw (:501-517), Config flips (:528-535), theta0/theta1 Beta
//            draws (:563-571), logLik trace with the OLD theta (:574-577 -> :734-752).
//
// Math is fp64 throughout (the tolerance in BASELINE.json is 1e-6 on probabilities whose logits are
// sums of ~1e3 terms of magnitude ~10); HBM traffic per sweep is 6 bytes per covered entry per pass.
#include "internal.h"
#include "cells.cuh"
#include <cstdlib>
#include <vector>

namespace cdl {

namespace {

constexpr int CELLS_THREADS = 512;
constexpr int STATS_THREADS = 512;
constexpr int STATS_WARPS = STATS_THREADS / 32;
constexpr int TABLE_THREADS = 256;

// One group of G lanes per cell (G >= KP); 32/G cells per warp at a time.  Each lane streams its share of
// the cell's 8-entry groups with a one-group software prefetch that also runs ACROSS cells, so the loads
// of the next cell are in flight during the reduction / softmax / draw of the current one.
template <int KP, int G, typename VIdx, bool SMEM>
__global__ void __launch_bounds__(CELLS_THREADS, (KP <= 8 ? 2 : 1)) k_cells(const CellsParams P) {
    constexpr int LOG = Log2<KP>::v;
    constexpr int LOGG = Log2<G>::v;
    constexpr int SH = LOGG - LOG;       // log2(lanes per clone) after the transposing reduction
    constexpr int CPW = 32 / G;          // cells per warp pass
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const double l0 = P.scal[SC_L0], l0c = P.scal[SC_L0C];
    uint32_t tab_sa = 0, msk_sa = 0;
    const double2* tab_g = nullptr;
    const uint32_t* msk_g = P.mask;
    if (SMEM) {
        // stage the per-variant weight table: (log th1_i - log th0, log(1-th1_i) - log(1-th0)) and the
        // Config_new row as a clone bit mask.  logLik[j,k] = base_j + sum_i C_ik * delta_ij; base_j
        // cancels in the softmax.
        double2* s_tab = reinterpret_cast<double2*>(smem_raw);
        uint32_t* s_msk = reinterpret_cast<uint32_t*>(smem_raw + sizeof(double2) * (size_t)P.N);
        for (int64_t i = threadIdx.x; i < P.N; i += blockDim.x) {
            const int64_t t = P.n_tab == 1 ? 0 : i;
            s_tab[i] = make_double2(P.l1[t] - l0, P.l1c[t] - l0c);
            s_msk[i] = P.mask[i];
        }
        __syncthreads();
        tab_sa = (uint32_t)__cvta_generic_to_shared(s_tab);
        msk_sa = (uint32_t)__cvta_generic_to_shared(s_msk);
    } else {
        tab_g = reinterpret_cast<const double2*>(P.l1);   // pre-built (du, dv) table in global memory
    }
    const int zero = P.zero;
    const int lane = threadIdx.x & 31;
    const int lig = lane & (G - 1);               // lane in group
    const int grp = lane >> LOGG;                 // group in warp
    const int gbase = lane & ~(G - 1);            // first lane of my group
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int K = P.K;
    const int k = lig >> SH;
    const bool rep = (lig & ((1 << SH) - 1)) == 0;
    const double my_logpsi = (k < K) ? P.logPsi[k] : 0.0;
    const uint4* a8 = reinterpret_cast<const uint4*>(P.ca);
    const uint4* b8 = reinterpret_cast<const uint4*>(P.cb);

    int64_t jb = warp * CPW;
    if (jb >= P.M) return;
    int64_t j = jb + grp;
    int64_t g0 = 0, g1 = 0;
    if (j < P.M) { g0 = P.rowgrp[j]; g1 = P.rowgrp[j + 1]; }
    Grp<VIdx> cur;
    cur.v = cur.a = cur.b = make_uint4(0, 0, 0, 0);
    bool have = (g0 + lig) < g1;
    if (have) cur.load(P.vidx, a8, b8, g0 + lig);

    for (; jb < P.M; jb += nwarps * CPW) {
        j = jb + grp;
        const bool valid = j < P.M;
        // row bounds of my group's NEXT cell, needed for the cross-cell prefetch
        const int64_t jn = j + nwarps * CPW;
        int64_t ng0 = 0, ng1 = 0;
        if (jn < P.M) { ng0 = P.rowgrp[jn]; ng1 = P.rowgrp[jn + 1]; }
        double acc[KP];
#pragma unroll
        for (int q = 0; q < KP; ++q) acc[q] = 0.0;
        int64_t g = g0 + lig;
        bool nhave;
        for (;;) {
            const int64_t gn = g + G;
            const bool more = gn < g1;
            Grp<VIdx> nxt;
            nxt.v = nxt.a = nxt.b = make_uint4(0, 0, 0, 0);
            nhave = more ? true : ((ng0 + lig) < ng1);
            if (nhave) nxt.load(P.vidx, a8, b8, more ? gn : ng0 + lig);
            if (have) accumulate_group<KP, VIdx, SMEM>(acc, cur, tab_sa, msk_sa, tab_g, msk_g, zero);
            cur = nxt;          // after the compute: the copy waits for the load, which had the whole
            have = nhave;       // accumulate_group (and, on the last pass, the epilogue) to land
            if (!more) break;
            g = gn;
        }
        g0 = ng0; g1 = ng1;
        // transposing butterfly inside the group: afterwards a lane holds the full sum of clone k
#pragma unroll
        for (int s = 0; s < LOG; ++s) {
            const int off = (G / 2) >> s;
            const int n = KP >> (s + 1);
            const bool up = (lig & off) != 0;
#pragma unroll
            for (int q = 0; q < KP / 2; ++q) {
                if (q < n) {
                    const double send = up ? acc[q] : acc[q + n];
                    const double keep = up ? acc[q + n] : acc[q];
                    acc[q] = keep + shfl_xor_d(send, off);
                }
            }
        }
        double L = acc[0];
#pragma unroll
        for (int off = (1 << SH) >> 1; off >= 1; off >>= 1) L += shfl_xor_d(L, off);
        L = (k < K) ? L + my_logpsi : -INFINITY;
        double mx = L;
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) mx = fmax(mx, shfl_xor_d(mx, off));
        const double ex = exp(L - mx);
        double sm = ex;
#pragma unroll
        for (int off = G / 2; off >= (1 << SH); off >>= 1) sm += shfl_xor_d(sm, off);
        const double pr = ex / sm;
        if (valid && rep && k < K) {
            if (P.prob_row) P.prob_row[j * K + k] = pr;
            if (P.prob_acc) P.prob_acc[j * K + k] += pr;
            if (P.prob_sq) P.prob_sq[j * K + k] += pr * pr;
        }
        // ---- sample(seq_len(K), 1, prob = pr): walk the DESCENDING-sorted CDF with one uniform ----
        const uint32_t gcell = (uint32_t)(P.cell_offset + j);
        double u = 0.5;
        if (valid) u = P.rp_u ? P.rp_u[j] : philox_uniform(P.key, P.chain, P.it, STREAM_ASSIGN, gcell);
        // shortcut: the first element of the sorted CDF is the maximum; if it is unique and u <= p_max the
        // walk stops there (confident cells, the common case)
        const unsigned top = __ballot_sync(0xffffffffu, rep && k < K && L == mx) & gmask;
        unsigned sel;
        const bool quick = (__popc(top) == 1) && (u * sm <= 1.0) && (u <= 1.0 / sm);
        if (__all_sync(0xffffffffu, quick)) {
            sel = (unsigned)(((__ffs(top) - 1) - gbase) >> SH);
        } else {
            double cum = 0.0;
            int rank = 0;
            bool tie = false;
#pragma unroll
            for (int l = 0; l < KP; ++l) {
                const double pl = shfl_d(pr, gbase + (l << SH));
                if (l < K) {
                    const bool before = (pl > pr) || (pl == pr && l < k);
                    if (before) { cum += pl; ++rank; }
                    tie |= (pl == pr) && (l != k) && (pr > 0.0);
                }
            }
            const bool cand = u <= cum + pr;
            const unsigned keyv = (k < K && (cand || rank == K - 1)) ? ((unsigned)rank << 8 | (unsigned)k)
                                                                     : 0xffffffffu;
            sel = __reduce_min_sync(gmask, keyv) & 0xffu;
            // the order inside a group of exactly tied probabilities is revsort's; it only matters when
            // the walk stops inside such a group
            const int sel_tied = __shfl_sync(0xffffffffu, (int)tie, gbase + ((int)sel << SH));
            if (sel_tied) {
                double pa[KP];
                int perm[KP];
#pragma unroll
                for (int l = 0; l < KP; ++l) {
                    pa[l] = __shfl_sync(gmask, pr, gbase + (l << SH));
                    perm[l] = l;
                }
                double tot = 0.0;
                for (int l = 0; l < K; ++l) tot += pa[l];
                for (int l = 0; l < K; ++l) pa[l] /= tot;
                revsort_dev(pa, perm, K);
                for (int l = 1; l < K; ++l) pa[l] += pa[l - 1];
                int jj = 0;
                while (jj < K - 1 && !(u <= pa[jj])) ++jj;
                sel = (unsigned)perm[jj];
            }
        }
        store_label(P, valid && lig == 0, j, (int)sel);
    }
}

// (du, dv) table in global memory for donors whose table does not fit shared memory
__global__ void k_build_tab(int64_t N, int n_tab, const double* __restrict__ l1, const double* __restrict__ l1c,
                            const double* __restrict__ scal, double2* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t t = n_tab == 1 ? 0 : i;
    out[i] = make_double2(l1[t] - scal[SC_L0], l1c[t] - scal[SC_L0C]);
}

// ------------------------------------------------------------------------------------------------
// "This rank's statistic table of sweep `epoch` is complete": release-store of the sweep number into every peer's flag
// array (cell-sharded runs over peer memory; n_peer = 0 otherwise).  The table was written with device-scope atomics
// into this GPU's memory and peers read it through NVLink from this GPU's L2, the coherence point, so a system-scope
// fence + release store is sufficient.  Issued by the LAST CTA of whichever kernel completed the table (the last CTA
// has seen every other CTA's ticket, each taken after a device-scope fence) -- no separate signalling kernel.
struct PeerSignal {
    int n_peer, my_rank;
    unsigned long long* peer_flags[MAX_PEERS];   // flag arrays in the peers' memory ([world] each, P2PState::my_flags)
    unsigned long long epoch;
};
// `changed`: this rank's table differs from the one of the previous sweep (a full pass, or a correction that moved at least
// one label); the flag carries 2 * epoch + changed.  A changed table is published with a release store (a system-scope
// fence: ~2 us on the sweep's critical path); a rank whose labels did not move has nothing to publish and sends the flag
// as a relaxed store.
__device__ __forceinline__ void signal_peers(const PeerSignal& S, unsigned int changed) {
    if ((int)threadIdx.x < S.n_peer) {
        unsigned long long* f = S.peer_flags[threadIdx.x] + S.my_rank;
        if (changed) st_release_sys_u64(f, 2ull * S.epoch + 1ull);
        else st_relaxed_sys_u64(f, 2ull * S.epoch);
    }
}
// bounded wait for a flag that a peer writes into this rank's memory (a peer that died must not hang this GPU)
__device__ __forceinline__ unsigned long long wait_flag(const unsigned long long* f, unsigned long long at_least,
                                                        unsigned int* error) {
    const long long t0 = clock64();
    unsigned long long v;
    while ((v = ld_acquire_sys_u64(f)) < at_least) {
        if (clock64() - t0 > 20000000000ll) { if (error) atomicExch(error, 1u); break; }   // ~10 s
        __nanosleep(100);
    }
    return v;
}
// every CTA calls this once, after its last write to the table; true in the CTA that finishes last
__device__ __forceinline__ bool last_cta_done(unsigned int* ticket) {
    __shared__ bool s_last;
    __threadfence();          // EVERY thread: its reductions (fire-and-forget RED) must have been performed at L2 before
    __syncthreads();          // the CTA takes its ticket -- a barrier alone does not wait for outstanding REDs
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) { __threadfence(); *ticket = 0u; }
    }
    __syncthreads();
    return s_last;
}

struct StatsParams {
    PeerSignal sig;
    unsigned int* done_ticket;      // zero between launches
    const uint32_t* seggrp;
    const uint16_t* flush;
    const uint16_t* vc;
    const uint16_t* va;
    const uint16_t* vb;
    const uint8_t* z;
    unsigned long long* sab;
    int64_t N, M;
    int K;
    int n_cb;             // cell blocks
    // incremental statistics (else ctl == nullptr): the same launch first tries to CORRECT the previous sweep's table
    // from the cells whose label changed (rows layout below); only above the threshold does it run the full pass --
    // into `zbuf`, a table that is all zero at rest, because the carried table cannot be cleared grid-wide first
    unsigned int* ctl;              // [0] changed count, [1] 1 if the last sweep took the full pass (2: corrected a changed table), [2] force-full flag, [3] full passes
    unsigned int threshold;
    unsigned long long* zbuf;       // [N*K]
    const uint32_t* rowgrp; const void* vidx; const uint16_t* ca; const uint16_t* cb; int vidx32;
    const uint32_t* chg_list; const uint8_t* chg_old;
    const int* plan;      // [grid + 2 * n_cb]: home block of each CTA, first CTA and CTA count of each block
    int chunks_per_warp;  // k_stats_units: work units per segment (k_stats_stream takes one static range per warp)
    unsigned int* work;   // [n_cb] unit counters of k_stats_units; zero on entry (k_table's last block re-arms them)
    unsigned long long* tl;   // device timeline or nullptr
};


// SA / SB are integer sums over the cells of each label, so when few labels changed in a sweep the table of the
// previous sweep is corrected instead of recomputed: for every changed cell (list built by the cell kernel, one
// atomic per warp), move its row's counts from the old label's column to the new one.  Exact (integers): the chain is
// bit-identical to the full pass.  Returns true if the table is done (the caller returns), false if the full pass has
// to run -- then into P.zbuf.
__device__ __forceinline__ bool stats_try_incremental(const StatsParams& P) {
    const unsigned int count = P.ctl[0];
    if (P.ctl[2] != 0 || count > P.threshold) return false;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // One work item = 32 consecutive entries of a changed cell's row, one per lane: a row of ~500 entries is spread over
    // 16 warps instead of being walked by one (its 16 dependent trips were ~12 us of pure latency on the sweep's
    // critical path whenever a single label had moved).
    constexpr int CHUNKS = 32;                                   // chunks per cell and round: rows up to 1024 entries in one
    const int64_t items = (int64_t)count * CHUNKS;
    for (int64_t it = warp; it < items; it += nwarps) {
        const int64_t q = it / CHUNKS;
        const int c = (int)(it - q * CHUNKS);
        const int64_t j = P.chg_list[q];
        const int64_t e0 = (int64_t)P.rowgrp[j] * GROUP, e1 = (int64_t)P.rowgrp[j + 1] * GROUP;
        if (e0 + (int64_t)c * 32 >= e1) continue;
        const int zo = P.chg_old[q], zn = P.z[j];
        for (int64_t e = e0 + (int64_t)c * 32 + lane; e < e1; e += 32 * CHUNKS) {
            const unsigned long long a = P.ca[e], b = P.cb[e];
            if ((a | b) == 0) continue;
            const int64_t v = P.vidx32 ? reinterpret_cast<const uint32_t*>(P.vidx)[e]
                                       : reinterpret_cast<const uint16_t*>(P.vidx)[e];
            const unsigned long long w = a | (b << 32);
            atomicAdd(&P.sab[v * P.K + zn], w);
            atomicAdd(&P.sab[v * P.K + zo], 0ull - w);     // the packed halves borrow and repay: net exact
        }
    }
    if (last_cta_done(P.done_ticket)) {
        // every other CTA has read the control words and finished its updates
        // ctl[1]: this sweep's table differs from the last one (the table kernel reads it: the carried sum is stale)
        if (threadIdx.x == 0) { P.ctl[0] = 0u; P.ctl[1] = count > 0 ? 2u : 0u; P.ctl[2] = 0u; }
        signal_peers(P.sig, count > 0 ? 1u : 0u);
    }
    tl_stamp(P.tl, 1, 3);
    return true;
}
// end of the full pass: the CTA that finishes last hands the table over
__device__ __forceinline__ void stats_finish_full(const StatsParams& P) {
    tl_stamp(P.tl, 1, 3);
    if (!P.ctl && P.sig.n_peer == 0) return;
    if (!last_cta_done(P.done_ticket)) return;
    if (P.ctl) {
        const int64_t NK = P.N * P.K;
        for (int64_t i = threadIdx.x; i < NK; i += blockDim.x) {       // the table moves to where the sweep expects it;
            P.sab[i] = P.zbuf[i];                                      // zbuf is all zero again for the next full pass
            P.zbuf[i] = 0ull;
        }
        if (threadIdx.x == 0) { P.ctl[0] = 0u; P.ctl[1] = 1u; P.ctl[2] = 0u; P.ctl[3] += 1u; }
        __threadfence();          // every thread's part of the copy is performed before anyone signals
        __syncthreads();
    }
    signal_peers(P.sig, 1u);
}

// One 8-entry group of a variant segment into the lane-private clone bins.  A and B share one 32-bit bin
// (A low half, B high half): one 32-bit shared load + store per entry.  The caller flushes the bins before
// either half can overflow (Donor::v_flush).
__device__ __forceinline__ void stats_group(const uint4 cw, const uint4 aw, const uint4 bw,
                                            const uint8_t* __restrict__ s_z, uint32_t* __restrict__ bins_lane) {
    const uint32_t cc[4] = {cw.x, cw.y, cw.z, cw.w};
    const uint32_t aa[4] = {aw.x, aw.y, aw.z, aw.w};
    const uint32_t bb[4] = {bw.x, bw.y, bw.z, bw.w};
    uint32_t zz[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t c = (e & 1) ? (cc[e >> 1] >> 16) : (cc[e >> 1] & 0xffffu);
        zz[e] = s_z[c];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t ab = __byte_perm(aa[e >> 1], bb[e >> 1], (e & 1) ? 0x7632 : 0x5410);   // a | b << 16
        bins_lane[zz[e] * 32] += ab;
    }
}

// column sums of the lane-private packed bins -> global table, then clear the bins
template <int KP>
__device__ __forceinline__ void stats_flush(uint32_t* __restrict__ bins, int lane, int K,
                                            unsigned long long* __restrict__ sab_row) {
    __syncwarp();
    // lane handles clone kk = lane % KP over a KP-wide slice of lanes, rotated so that the lanes of a
    // wavefront touch distinct banks
    const int kk = lane & (KP - 1);
    const int h = lane / KP;                 // 0 .. 32/KP-1
    unsigned long long sa = 0, sb = 0;
#pragma unroll
    for (int s = 0; s < KP; ++s) {
        const int t = h * KP + ((s + kk) & (KP - 1));
        const uint32_t v = bins[kk * 32 + t];
        sa += v & 0xffffu;
        sb += v >> 16;
    }
#pragma unroll
    for (int off = KP; off < 32; off <<= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, off);
        sb += __shfl_xor_sync(0xffffffffu, sb, off);
    }
    if (lane < K && (sa | sb)) atomicAdd(&sab_row[lane], sa | (sb << 32));
    __syncwarp();
#pragma unroll
    for (int q = 0; q < KP; ++q) bins[q * 32 + lane] = 0u;
    __syncwarp();
}

// Every CTA stages the labels of ONE cell block in shared memory (CTAs are dealt to the cell blocks in
// proportion to their entries; with more cell blocks than CTAs a CTA walks several).  The block's variant
// segments are contiguous in memory, so the block's groups are cut into one contiguous range per warp and each
// warp STREAMS its range: the loads of the next 32 groups are always in flight while the current 32 are binned,
// across segment boundaries too, and the next 32 segment boundaries / flush intervals sit in one register per
// lane (shuffled out as needed) -- no per-unit atomics or dependent metadata loads, which on a 1/8 shard cost
// as much as the work itself.  A range may start or end inside a segment: the partial sums simply meet in the
// global table (integer atomics, exact).  A lane keeps private per-clone bins in shared memory (A and B packed
// 16 + 16 bits, flushed before either half can overflow).  32 warps per SM (2 CTAs x 16 warps, <= 64 registers):
// the kernel is bound by the latency of the shared-memory update chain as much as by HBM.
template <int KP>
__global__ void __launch_bounds__(STATS_THREADS, 2) k_stats_stream(const StatsParams P) {
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bool waited = false;
    unsigned long long* const sab = P.ctl ? P.zbuf : P.sab;      // where a full pass accumulates
    tl_stamp(P.tl, 1, 0);
    if (P.ctl) {
        // incremental statistics: the control words come from the cell kernel, so wait first; most sweeps end here
        pdl_wait();
        tl_stamp(P.tl, 1, 1);
        waited = true;
        if (stats_try_incremental(P)) return;
    }
    uint32_t* s_bins = reinterpret_cast<uint32_t*>(smem_raw);                  // [WARPS][KP][32]
    uint8_t* s_z = smem_raw + sizeof(uint32_t) * STATS_WARPS * KP * 32;        // [cells of the staged block]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t* bins = s_bins + (size_t)w * KP * 32;
    const uint4* c8 = reinterpret_cast<const uint4*>(P.vc);
    const uint4* a8 = reinterpret_cast<const uint4*>(P.va);
    const uint4* b8 = reinterpret_cast<const uint4*>(P.vb);
#pragma unroll
    for (int q = 0; q < KP; ++q) bins[q * 32 + lane] = 0u;
    const int G = (int)gridDim.x;
    int cb, cb_step;
    int64_t r, R;            // this warp's rank among the R warps that share the cell block
    if (P.n_cb <= G) {
        cb = P.plan[blockIdx.x];                                  // home block of this CTA
        const int first = P.plan[G + cb], count = P.plan[G + P.n_cb + cb];
        r = (int64_t)((int)blockIdx.x - first) * STATS_WARPS + w;
        R = (int64_t)count * STATS_WARPS;
        cb_step = P.n_cb;
    } else {
        cb = (int)blockIdx.x; cb_step = G; r = w; R = STATS_WARPS;
    }
    // Everything up to pdl_wait() reads only the donor's layout (never written after the load), so under a
    // programmatic dependent launch it runs while the cell kernel is still draining: the warp's range, the
    // search for its first segment and the loads of its first 32 groups.
    for (; cb < P.n_cb; cb += cb_step) {
        const uint32_t* sg = P.seggrp + (int64_t)cb * P.N;       // sg[0..N]: group offsets of the block's segments
        const uint16_t* fl = P.flush + (int64_t)cb * P.N;
        const int64_t G0 = sg[0], G1 = sg[P.N];
        // one static contiguous range of the block's groups per warp
        const int64_t lo = G0 + ((G1 - G0) * r) / R, hi = G0 + ((G1 - G0) * (r + 1)) / R;
        const bool has = lo < hi;
        int64_t i = 0, win = 0, cur = lo, send = hi;
        uint32_t wend = 0, wfl = 0;
        int flush_every = 0;
        uint4 cw = make_uint4(0, 0, 0, 0), aw = cw, bw = cw;
        auto seg_end_of = [&](int64_t seg) -> int64_t {
            return (int64_t)__shfl_sync(0xffffffffu, wend, (int)(seg - win));
        };
        if (has) {
            // segment that holds group `lo`: largest i with sg[i] <= lo (32-ary search, the lanes probe in parallel)
            int64_t span = P.N;                 // answer in [i, i + span)
            while (span > 1) {
                const int64_t stepw = (span + 31) / 32;
                const int64_t probe = min(i + stepw * lane, P.N - 1);
                const bool le = (int64_t)sg[probe] <= lo && (i + stepw * lane) < i + span;
                const unsigned m = __ballot_sync(0xffffffffu, le);
                const int top = 31 - __clz((int)m);                // lane 0 always votes (sg[i] <= lo)
                const int64_t ni = i + stepw * top;
                span = min(stepw, i + span - ni);
                i = ni;
            }
            // boundary window: lane l holds sg[win + 1 + l] (end of segment win + l) and fl[win + l]
            win = i;
            wend = sg[min(win + 1 + lane, P.N)];
            wfl = fl[min(win + lane, P.N - 1)];
            send = min(seg_end_of(i), hi);
            while (send <= cur) {              // cannot happen for i found above, but keep the invariant explicit
                ++i;
                if (i - win >= 32) { win = i; wend = sg[min(win + 1 + lane, P.N)]; wfl = fl[min(win + lane, P.N - 1)]; }
                send = min(seg_end_of(i), hi);
            }
            flush_every = (int)__shfl_sync(0xffffffffu, wfl, (int)(i - win));
            const int64_t gi = min(cur + lane, send - 1);
            cw = ld_stream(c8 + gi); aw = ld_stream(a8 + gi); bw = ld_stream(b8 + gi);
        }
        if (!waited) {
            pdl_wait();
            waited = true;
        }
        __syncthreads();                       // all warps are done with the previous block's labels
        {
            const int64_t c0 = (int64_t)cb * CELL_BLOCK;
            const int ncell = (int)min((int64_t)CELL_BLOCK, P.M - c0);
            const int nvec = ncell >> 4;
            const uint4* src = reinterpret_cast<const uint4*>(P.z + c0);
            uint4* dst = reinterpret_cast<uint4*>(s_z);
            for (int t = threadIdx.x; t < nvec; t += blockDim.x) dst[t] = src[t];
            for (int t = (nvec << 4) + threadIdx.x; t < ncell; t += blockDim.x) s_z[t] = P.z[c0 + t];
        }
        __syncthreads();
        if (!has) continue;
        int since = 0;
        while (cur < hi) {
            // where the NEXT iteration starts: the next 32 groups of this segment, or the start of the next
            // non-empty segment
            int64_t nbase, nsend = send, ni = i;
            int nfl = flush_every;
            const bool seg_done = cur + 32 >= send;
            if (!seg_done) nbase = cur + 32;
            else {
                nbase = send;
                if (nbase < hi) {
                    do {
                        ++ni;
                        if (ni - win >= 32) { win = ni; wend = sg[min(win + 1 + lane, P.N)]; wfl = fl[min(win + lane, P.N - 1)]; }
                        nsend = min(seg_end_of(ni), hi);
                    } while (nsend <= nbase);
                    nfl = (int)__shfl_sync(0xffffffffu, wfl, (int)(ni - win));
                }
            }
            uint4 cn = cw, an = aw, bn = bw;
            if (nbase < hi) {
                const int64_t gi = min(nbase + lane, nsend - 1);
                cn = ld_stream(c8 + gi); an = ld_stream(a8 + gi); bn = ld_stream(b8 + gi);
            }
            unsigned long long* sab_row = sab + i * P.K;
            if (cur + lane < send) {
                if (flush_every) stats_group(cw, aw, bw, s_z, bins + lane);
                else {
                    // depths above 8191 in this segment: a single group could overflow a 16-bit half; take the
                    // (rare) direct path, one 64-bit atomic per entry
                    const uint32_t cc[4] = {cw.x, cw.y, cw.z, cw.w};
                    const uint32_t aa[4] = {aw.x, aw.y, aw.z, aw.w};
                    const uint32_t bb[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t c = (e & 1) ? (cc[e >> 1] >> 16) : (cc[e >> 1] & 0xffffu);
                        const unsigned long long a = (e & 1) ? (aa[e >> 1] >> 16) : (aa[e >> 1] & 0xffffu);
                        const unsigned long long b = (e & 1) ? (bb[e >> 1] >> 16) : (bb[e >> 1] & 0xffffu);
                        if (a | b) atomicAdd(&sab_row[s_z[c]], a | (b << 32));
                    }
                }
            }
            ++since;
            if (flush_every && (seg_done || since == flush_every)) { stats_flush<KP>(bins, lane, P.K, sab_row); since = 0; }
            if (seg_done) since = 0;
            cw = cn; aw = an; bw = bn;
            cur = nbase; send = nsend; i = ni; flush_every = nfl;
        }
    }
    if (!waited) pdl_wait();
    stats_finish_full(P);
}

// The same statistics for LARGE shards (many segments per warp): the warps pull whole variant segments (or
// 1/parts of one) from the cell block's counter, each warp on its own.  One atomic and two dependent metadata
// loads per unit are noise when a unit is a few thousand entries and 30+ units are queued per warp, and the
// per-iteration bookkeeping is lighter than the streaming kernel's (C4 on one GPU: 0.51 ms vs 0.58 ms); on a
// 1/8 shard or at C3 the streaming kernel wins by 15-25 %.  A CTA starts on its home block and then walks the
// others, helping wherever units are left.
template <int KP>
__global__ void __launch_bounds__(STATS_THREADS, 2) k_stats_units(const StatsParams P) {
    pdl_trigger();
    tl_stamp(P.tl, 1, 0);
    pdl_wait();
    tl_stamp(P.tl, 1, 1);
    if (P.ctl && stats_try_incremental(P)) return;   // incremental statistics: most sweeps end here
    unsigned long long* const sab = P.ctl ? P.zbuf : P.sab;      // where a full pass accumulates
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_bins = reinterpret_cast<uint32_t*>(smem_raw);                  // [WARPS][KP][32]
    uint8_t* s_z = smem_raw + sizeof(uint32_t) * STATS_WARPS * KP * 32;        // [cells of the staged block]
    __shared__ unsigned int s_peek;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t* bins = s_bins + (size_t)w * KP * 32;
    const uint4* c8 = reinterpret_cast<const uint4*>(P.vc);
    const uint4* a8 = reinterpret_cast<const uint4*>(P.va);
    const uint4* b8 = reinterpret_cast<const uint4*>(P.vb);
#pragma unroll
    for (int q = 0; q < KP; ++q) bins[q * 32 + lane] = 0u;
    const int G = (int)gridDim.x;
    const int cb0 = P.n_cb <= G ? P.plan[blockIdx.x] : (int)(blockIdx.x % (unsigned int)P.n_cb);
    const int parts = P.chunks_per_warp;                 // here: work units per segment
    const unsigned int n_units = (unsigned int)(P.N * parts);
    for (int step = 0; step < P.n_cb; ++step) {
        int cb = cb0 + step;
        if (cb >= P.n_cb) cb -= P.n_cb;
        __syncthreads();                       // all warps are done with the previous block's labels / s_peek
        if (threadIdx.x == 0) s_peek = *((volatile unsigned int*)(P.work + cb));
        __syncthreads();
        if (s_peek >= n_units) continue;       // nothing left in that block
        {
            const int64_t c0 = (int64_t)cb * CELL_BLOCK;
            const int ncell = (int)min((int64_t)CELL_BLOCK, P.M - c0);
            const int nvec = ncell >> 4;
            const uint4* src = reinterpret_cast<const uint4*>(P.z + c0);
            uint4* dst = reinterpret_cast<uint4*>(s_z);
            for (int t = threadIdx.x; t < nvec; t += blockDim.x) dst[t] = src[t];
            for (int t = (nvec << 4) + threadIdx.x; t < ncell; t += blockDim.x) s_z[t] = P.z[c0 + t];
        }
        __syncthreads();
        for (;;) {
            unsigned int u = 0;
            if (lane == 0) u = atomicAdd(P.work + cb, 1u);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= n_units) break;
            const int64_t i = (int64_t)(u / (unsigned int)parts);
            const int part = (int)(u - (unsigned int)i * (unsigned int)parts);
            const int64_t seg = (int64_t)cb * P.N + i;
            int64_t g0 = P.seggrp[seg], g1 = P.seggrp[seg + 1];
            if (parts > 1) {
                const int64_t len = g1 - g0;
                g1 = g0 + (len * (part + 1)) / parts;
                g0 = g0 + (len * part) / parts;
            }
            if (g0 >= g1) continue;
            const int flush_every = P.flush[seg];
            unsigned long long* sab_row = sab + i * P.K;
            if (flush_every == 0) {
                for (int64_t g = g0 + lane; g < g1; g += 32) {
                    const uint4 cw = ld_stream(c8 + g), aw = ld_stream(a8 + g), bw = ld_stream(b8 + g);
                    const uint32_t cc[4] = {cw.x, cw.y, cw.z, cw.w};
                    const uint32_t aa[4] = {aw.x, aw.y, aw.z, aw.w};
                    const uint32_t bb[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t c = (e & 1) ? (cc[e >> 1] >> 16) : (cc[e >> 1] & 0xffffu);
                        const unsigned long long a = (e & 1) ? (aa[e >> 1] >> 16) : (aa[e >> 1] & 0xffffu);
                        const unsigned long long b = (e & 1) ? (bb[e >> 1] >> 16) : (bb[e >> 1] & 0xffffu);
                        if (a | b) atomicAdd(&sab_row[s_z[c]], a | (b << 32));
                    }
                }
                continue;
            }
            int64_t g = g0 + lane;
            int64_t gi = min(g, g1 - 1);
            uint4 cw = ld_stream(c8 + gi), aw = ld_stream(a8 + gi), bw = ld_stream(b8 + gi);
            int since = 0;
            for (int64_t gb = g0; gb < g1; gb += 32, g += 32) {       // warp-uniform trip count
                gi = min(g + 32, g1 - 1);
                const uint4 cn = ld_stream(c8 + gi), an = ld_stream(a8 + gi), bn = ld_stream(b8 + gi);
                if (g < g1) stats_group(cw, aw, bw, s_z, bins + lane);
                cw = cn; aw = an; bw = bn;
                if (++since == flush_every) { stats_flush<KP>(bins, lane, P.K, sab_row); since = 0; }
            }
            if (since) stats_flush<KP>(bins, lane, P.K, sab_row);
        }
    }
    stats_finish_full(P);
}

// ------------------------------------------------------------------------------------------------
struct TableParams {
    int64_t N;
    int K, wise, relax, keep_base, learn_relax_next;
    double prior0[2], prior1[2], relax_prior[2];
    const double* prior1_mat;  // [n_el*2] or nullptr
    int64_t n_el;
    const uint32_t* cfg0;
    uint32_t* mask;
    const unsigned long long* sab;
    unsigned long long* sab_next;   // table the NEXT sweep's k_stats accumulates into: cleared here, element by
                                    // element, once this sweep's value has been read (saves a memset per sweep);
                                    // nullptr with incremental statistics on one GPU, which carry the table over in place
    int carry_local;                // incremental statistics over peer memory: sab_next (the other half of the arena)
                                    // receives THIS rank's table instead of zeros, for the next sweep to correct
    unsigned long long* dbg_s34;    // test hook: [2N] the integers added to the theta1 priors (S3_i, S4_i), else nullptr
    unsigned long long* dbg_tot;    // test hook: [5] S1, S2, S3, S4, diff1 totals
    unsigned int* sell_work;        // k_cells_sell's work counter, re-armed by the last block
    unsigned int* stats_work;       // k_stats_units' per-cell-block unit counters [n_cb], re-armed likewise
    int n_cb;
    const double* sd1;         // wise = element: SD[i,k] = sum_{j: z_j = k} A ln th1_ij + B ln(1 - th1_ij), else nullptr
    double* scal;
    double* l1;
    double* l1c;
    double W_log;
    double* theta0_all; double* theta1_all; double* loglik_all; double* relax_all; uint8_t* config_all;
    int64_t T;
    uint32_t it;               // 1-based row being filled
    uint32_t rng_it;           // iteration in the Philox keys
    PhiloxKey key; uint32_t chain;
    const double* rp_u_config; const double* rp_theta0; const double* rp_theta1; const double* rp_relax_next;
    double* part_d; unsigned long long* part_u; unsigned int* ticket;
    // cell-sharded runs over peer memory (NVLink): the other ranks' statistic tables are summed on the fly
    // once their "table of sweep `epoch` is complete" flags have arrived (else n_peer = 0)
    int n_peer;
    const unsigned long long* peer_sab[MAX_PEERS];
    const unsigned long long* my_flags;       // [world] written by the peers into THIS rank's memory
    int peer_rank[MAX_PEERS];
    unsigned long long epoch;
    unsigned int* p2p_error;                  // set to 1 if a peer's flag did not arrive in time
    // Carried sum (incremental statistics on cell shards; nullptr = off): this rank's copy of the table summed over ALL
    // ranks, rewritten whenever any rank's table changed and read INSTEAD of the seven peers' tables (9 MB per rank and
    // sweep at C4 on 8 GPUs) when none did -- in a settled chain almost every sweep
    unsigned long long* g_own;
    const unsigned int* ctl;                  // chg_ctl: [1] != 0 if this rank's table changed in this sweep
    unsigned long long* tl;                   // device timeline or nullptr
    unsigned long long* tl_hist;
};

// Block totals of k_table's six per-thread terms.  On return sm_u[q * 32] (q = 0..4) and sm_d[0] hold them; every
// thread has passed one barrier after the warp partials were stored and one before they were overwritten.
__device__ __forceinline__ void table_block_sums(unsigned long long a0, unsigned long long a1, unsigned long long a2,
                                                 unsigned long long a3, unsigned long long a4, double d,
                                                 unsigned long long* sm_u, double* sm_d) {
    a0 = warp_sum_u64(a0); a1 = warp_sum_u64(a1); a2 = warp_sum_u64(a2); a3 = warp_sum_u64(a3); a4 = warp_sum_u64(a4);
    d = warp_sum_d(d);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (int)(blockDim.x >> 5);
    __syncthreads();                       // earlier readers of sm_u / sm_d are done
    if (lane == 0) {
        sm_u[w] = a0; sm_u[32 + w] = a1; sm_u[64 + w] = a2; sm_u[96 + w] = a3; sm_u[128 + w] = a4;
        sm_d[w] = d;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        unsigned long long r = 0;
        for (int q = 0; q < nw; ++q) r += sm_u[threadIdx.x * 32 + q];
        sm_u[threadIdx.x * 32] = r;
    } else if (threadIdx.x == 5) {
        double r = 0.0;
        for (int q = 0; q < nw; ++q) r += sm_d[q];          // warps in order, as everywhere: fixed rounding
        sm_d[0] = r;
    }
    if (threadIdx.x < 32) __syncwarp();
}

// Gamma jobs of a k_table block: two per variant of the block (theta1: X and Y of the Beta), then theta0 (2), the next
// relax rate (2) and the global theta1 (2), which only the block that finishes last carries out.  The Philox identity
// of job j -- false if this launch does not draw it -- depends on the launch parameters only.
template <int KP>
__host__ __device__ constexpr int table_njob() { return 2 * (TABLE_THREADS / KP) + 6; }
template <int KP>
__device__ __forceinline__ bool table_job(const TableParams& P, int j, uint32_t& stream, uint32_t& idx, uint32_t& it) {
    constexpr int VPB = TABLE_THREADS / KP;
    it = P.rng_it; idx = 0; stream = STREAM_THETA1;
    if (j < 2 * VPB) {
        const int64_t iv = (int64_t)blockIdx.x * VPB + (j >> 1);
        idx = (uint32_t)iv;
        return P.wise == 1 && iv < P.N && !P.rp_theta1;
    }
    const int q = j - 2 * VPB;
    if (q < 2) { stream = STREAM_THETA0; return !P.rp_theta0; }
    if (q < 4) { stream = STREAM_RELAX; it = P.rng_it + 1; return P.learn_relax_next && (int64_t)P.it < P.T && !P.rp_relax_next; }
    return q < 6 && P.wise == 0 && !P.rp_theta1;
}

// One thread per (variant, clone) element; the KP lanes of a variant sit in one warp.
//   phase A  Config flip of the element and its contribution to the theta statistics / logLik trace, reduced
//            over the variant's lanes with shuffles; block partials go through a ticket and the last block
//            reduces them in block order (deterministic).
//   phase B  every Beta draw of the sweep is two independent Gamma draws (Marsaglia-Tsang on their own Philox
//            sub-streams); each Gamma is one job on one thread: 2 per variant of the block for theta1, plus --
//            in the last block -- theta0, the next relax rate and the global theta1.  A serial fp64
//            rejection sampler is ~10 cycles per dependent instruction, so the kernel's duration is one Gamma
//            draw, not a chain of them.
template <int KP>
__global__ void __launch_bounds__(TABLE_THREADS, 5) k_table(const TableParams P) {
    constexpr int VPB = TABLE_THREADS / KP;          // variants per block
    constexpr int NJOB = table_njob<KP>();
    constexpr int NCAND = TABLE_THREADS / NJOB < 3 ? TABLE_THREADS / NJOB : 3;     // attempts prepared ahead per job
    static_assert(NJOB <= TABLE_THREADS, "one Gamma job per thread");
    __shared__ unsigned long long sm_u[5 * 32];
    __shared__ double sm_d[32];
    __shared__ unsigned long long s_s3[VPB], s_s4[VPB], s_tot[5];
    __shared__ double s_gam[NJOB];
    __shared__ double s_cx[NJOB * NCAND], s_cu[NJOB * NCAND];      // Gamma candidates: normal deviate, acceptance uniform
    __shared__ bool is_last;
    pdl_trigger();
    tl_stamp(P.tl, 2, 0);
    // The first NCAND attempts of every Gamma job of this block, as far as they do not depend on the shape (common.cuh:
    // gamma_candidate -- two Philox blocks, a logarithm, a square root and a cosine each), one (job, attempt) per thread.
    // They need nothing but the launch parameters, so they are made BEFORE the wait for the statistic kernel (and, on cell
    // shards, for the peers' tables): the sweep's critical path keeps only the acceptance tests.  The six global jobs
    // (theta0, next relax rate, global theta1) are prepared by every block, because which block finishes last is not
    // known yet.
    if ((int)threadIdx.x < NJOB * NCAND) {
        const int j = (int)threadIdx.x / NCAND, t = (int)threadIdx.x - j * NCAND;
        uint32_t stream, idx, it;
        if (table_job<KP>(P, j, stream, idx, it)) {
            PhiloxSeq sq(P.key, P.chain, it, stream, idx, ((uint32_t)(j & 1) << 31) + 2u * (uint32_t)t);
            double x, u3;
            gamma_candidate(sq, x, u3);
            s_cx[threadIdx.x] = x; s_cu[threadIdx.x] = u3;
        }
    }
    tl_stamp(P.tl, 2, 7);
    pdl_wait();
    tl_stamp(P.tl, 2, 1);
    const int k = threadIdx.x & (KP - 1);
    const int v = threadIdx.x / KP;
    const int64_t i = (int64_t)blockIdx.x * VPB + v;
    const double l0 = P.scal[SC_L0], l0c = P.scal[SC_L0C];
    const double odd1 = P.scal[SC_ODD1];         // Config == 1: log(prior) - log(1 - prior), prior = 1 - r: log(1 - r) - log(r)
    const double odd0 = -odd1;                   // Config == 0: prior = r: log(r) - log(1 - r), the exact negative
    unsigned long long S1 = 0, S2 = 0, S3 = 0, S4 = 0, diff1 = 0;
    double ll = 0.0;
    uint32_t bitk = 0;
    bool use_sum = false;                      // block-uniform: the carried sum is current, no peer table is read
    if (P.n_peer > 0) {
        // fused all-reduce: wait until every peer has published its table of this sweep (last CTA of the peer's statistic
        // kernel), then read the peers' tables straight over NVLink below -- or, when no rank's table changed in this
        // sweep, the carried sum in local memory.
        __shared__ unsigned int s_any_changed;
        if (threadIdx.x == 0) s_any_changed = P.g_own ? (P.ctl[1] != 0u ? 1u : 0u) : 1u;
        __syncthreads();
        if ((int)threadIdx.x < P.n_peer) {
            const unsigned long long v = wait_flag(P.my_flags + P.peer_rank[threadIdx.x], 2ull * P.epoch, P.p2p_error);
            if (v & 1ull) atomicOr(&s_any_changed, 1u);
        }
        __syncthreads();
        use_sum = s_any_changed == 0u;
        tl_stamp(P.tl, 2, 2);
    }
    if (i < P.N && k < P.K) {
        const bool element = P.sd1 != nullptr;
        const int64_t t = P.wise == 1 ? i : 0;
        const double l1 = element ? 0.0 : P.l1[t], l1c = element ? 0.0 : P.l1c[t];
        const double du = l1 - l0, dv = l1c - l0c;
        const uint32_t c0 = P.cfg0[i];
        unsigned long long pk = P.sab[i * P.K + k];
        const unsigned long long pk_local = pk;
        if (use_sum) pk = P.g_own[i * P.K + k];
        else if (P.n_peer > 0) {
            // all peers' loads in flight together: one NVLink round trip, not one per peer
            unsigned long long pv[MAX_PEERS];
#pragma unroll
            for (int q = 0; q < MAX_PEERS; ++q) pv[q] = q < P.n_peer ? ld_peer_u64(P.peer_sab[q] + i * P.K + k) : 0ull;
#pragma unroll
            for (int q = 0; q < MAX_PEERS; ++q) pk += pv[q];
            if (P.g_own) P.g_own[i * P.K + k] = pk;         // the carried sum starts over from this sweep's tables
        }
        if (P.sab_next) P.sab_next[i * P.K + k] = P.carry_local ? pk_local : 0ull;
        const double SA = (double)(uint32_t)pk, SB = (double)(uint32_t)(pk >> 32);
        const double sd = element ? P.sd1[i * P.K + k] : 0.0;       // per-entry theta1: the theta1 part, summed
        const uint32_t cbit = (c0 >> k) & 1u;
        uint32_t bit = cbit;
        if (P.relax) {
            double prior;
            if (P.keep_base && k == 0) prior = cbit ? INFINITY : -INFINITY;
            else prior = cbit ? odd1 : odd0;
            double x = (element ? sd - (SA * l0 + SB * l0c) : SA * du + SB * dv) + prior;
            // Beyond +-40 the draw does not depend on its uniform: e^x + 1 == e^x in double, so p == 1.0 and rbinom
            // returns 1 without a uniform (:533, nmath/rbinom.c); at the other end 1 - p == 1.0, every uniform is below it
            // and the draw is 0.  With tens of thousands of cells behind each variant nearly every element is there: no
            // exp, no division, no Philox block for it (counter-based streams: skipping a draw moves no other draw).
            if (x >= 40.0) bit = 1u;
            else if (x <= -40.0) bit = 0u;
            else {
                const double ex = exp(x);              // (the reference's clip to [-50, 50] cannot bind in here)
                const double p = ex / (ex + 1.0);
                const uint32_t idx = (uint32_t)(i + (int64_t)k * P.N);
                const double u = P.rp_u_config ? P.rp_u_config[idx]
                                               : philox_uniform(P.key, P.chain, P.rng_it, STREAM_CONFIG, idx);
                bit = (uint32_t)rbinom1(p, u);
            }
        }
        // Config_all row (:535); with relax off the row holds Config itself (the reference stops
        // with "object 'Config_new' not found" there)
        P.config_all[(int64_t)(P.it - 1) * P.N * P.K + i + (int64_t)k * P.N] = (uint8_t)bit;
        bitk = bit << k;
        if (bit) {
            S3 = (uint32_t)pk; S4 = (uint32_t)(pk >> 32);
            ll = element ? sd : SA * l1 + SB * l1c;
        } else {
            S1 = (uint32_t)pk; S2 = (uint32_t)(pk >> 32);
            ll = SA * l0 + SB * l0c;
        }
        if (k >= 1) diff1 = (bit != cbit);
    }
    // per-variant sums over the KP lanes (fixed butterfly order)
    unsigned long long v3 = S3, v4 = S4;
#pragma unroll
    for (int off = 1; off < KP; off <<= 1) {
        v3 += __shfl_xor_sync(0xffffffffu, v3, off);
        v4 += __shfl_xor_sync(0xffffffffu, v4, off);
        bitk |= __shfl_xor_sync(0xffffffffu, bitk, off);
    }
    if (k == 0) {
        s_s3[v] = v3; s_s4[v] = v4;
        if (i < P.N) P.mask[i] = bitk;
    }
    // block totals of the five integer sums and the logLik term: warp butterflies, then thread q adds component
    // q over the warps in order (one barrier; the double sum keeps a fixed order)
    table_block_sums(S1, S2, S3, S4, diff1, ll, sm_u, sm_d);          // (the barrier makes s_s3 / s_s4 visible)
    if (threadIdx.x < 6) {
        if (threadIdx.x < 5) P.part_u[(size_t)blockIdx.x * 8 + threadIdx.x] = sm_u[threadIdx.x * 32];
        else P.part_d[blockIdx.x] = sm_d[0];
        __threadfence();
    }
    if (threadIdx.x < 32) {
        __syncwarp();
        tl_stamp(P.tl, 2, 3);
        if (threadIdx.x == 0) {
            const unsigned int tk = atomicAdd(P.ticket, 1u);
            is_last = (tk == gridDim.x - 1);
        }
    }
    __syncthreads();
    const bool last = is_last;                        // block-uniform
    if (last) {
        __threadfence();
        // deterministic final reduction in block order
        unsigned long long t1 = 0, t2 = 0, t3 = 0, t4 = 0, td = 0;
        double tl = 0.0;
        // (L2 loads, all of a thread's rows in flight together: this stretch is on the critical path of every sweep)
#pragma unroll 4
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            const ulonglong2* pu = reinterpret_cast<const ulonglong2*>(P.part_u + (size_t)b * 8);
            const ulonglong2 x01 = __ldcg(pu), x23 = __ldcg(pu + 1);
            const unsigned long long x4 = __ldcg(P.part_u + (size_t)b * 8 + 4);
            const double xl = __ldcg(P.part_d + b);
            t1 += x01.x; t2 += x01.y; t3 += x23.x; t4 += x23.y; td += x4;
            tl += xl;
        }
        table_block_sums(t1, t2, t3, t4, td, tl, sm_u, sm_d);
        if (threadIdx.x < 5) {
            s_tot[threadIdx.x] = sm_u[threadIdx.x * 32];
            if (P.dbg_tot) P.dbg_tot[threadIdx.x] = sm_u[threadIdx.x * 32];
        }
        if (threadIdx.x == 5) P.loglik_all[(int64_t)P.it - 1] = sm_d[0] + P.W_log;
        if (threadIdx.x == 0) {
            *P.ticket = 0;
            if (P.sell_work) *P.sell_work = 0u;
        }
        if (P.stats_work)
            for (int b = threadIdx.x; b < P.n_cb; b += blockDim.x) P.stats_work[b] = 0u;
        __syncthreads();
        tl_stamp(P.tl, 2, 4);
    }
    // ---- phase B: one Gamma job per thread; the candidates were made before the wait, what is left is the shape ----
    {
        const int j = threadIdx.x;
        double shape = -1.0;
        uint32_t stream = 0, idx = 0, it = P.rng_it;
        if (j < NJOB && (j < 2 * VPB || last) && table_job<KP>(P, j, stream, idx, it)) {
            if (j < 2 * VPB) {
                const int64_t iv = (int64_t)blockIdx.x * VPB + (j >> 1);
                const double pa = P.prior1_mat ? P.prior1_mat[iv] : P.prior1[0];
                const double pb = P.prior1_mat ? P.prior1_mat[P.n_el + iv] : P.prior1[1];
                shape = (j & 1) ? pb + (double)s_s4[j >> 1] : pa + (double)s_s3[j >> 1];
                if (P.dbg_s34) P.dbg_s34[(j & 1) ? P.N + iv : iv] = (j & 1) ? s_s4[j >> 1] : s_s3[j >> 1];
            } else {
                const int q = j - 2 * VPB;               // 0,1 theta0; 2,3 next relax rate; 4,5 global theta1
                if (q < 2) shape = P.prior0[q] + (double)s_tot[q];
                else if (q < 4) {
                    const double d1 = (double)s_tot[4];
                    shape = (q == 2) ? P.relax_prior[0] + d1
                                     : P.relax_prior[1] + (double)P.N * (double)(P.K - 1) - d1;
                } else {
                    const double pa = P.prior1_mat ? P.prior1_mat[0] : P.prior1[0];
                    const double pb = P.prior1_mat ? P.prior1_mat[1] : P.prior1[1];
                    shape = (q == 4) ? pa + (double)s_tot[2] : pb + (double)s_tot[3];
                }
            }
        }
        if (shape >= 0.0) {
            const uint32_t sub0 = (uint32_t)(j & 1) << 31;              // X and Y: disjoint halves
            double g;
            if (shape < 1.0) {
                // the boost block comes first and shifts the attempts by one block: the serial sampler (rare: no read of
                // the variant supports the allele in any cell of the clones that carry it)
                PhiloxSeq sq(P.key, P.chain, it, stream, idx, sub0);
                g = gamma_mt(shape, sq);
            } else {
                const double d = shape - 1.0 / 3.0;
                const double c = 1.0 / sqrt(9.0 * d);
                bool done = false;
                double vv = 0.0;
                g = d;
#pragma unroll
                for (int t = 0; t < NCAND; ++t) {
                    if (!done && gamma_accept(d, c, s_cx[j * NCAND + t], s_cu[j * NCAND + t], vv)) { g = d * vv; done = true; }
                }
                if (!done) {                                            // (~1e-4 of the jobs with three candidates)
                    PhiloxSeq sq(P.key, P.chain, it, stream, idx, sub0 + 2u * (uint32_t)NCAND);
                    g = gamma_mt_attempts(d, c, sq, 1000);
                }
            }
            s_gam[j] = g;
        }
    }
    __syncthreads();
    tl_stamp(P.tl, 2, 5);
    if (P.wise == 1 && threadIdx.x < VPB) {
        const int64_t iv = (int64_t)blockIdx.x * VPB + threadIdx.x;
        if (iv < P.N) {
            const double th = P.rp_theta1 ? clamp_theta(P.rp_theta1[iv])
                                          : beta_from_gammas(s_gam[2 * threadIdx.x], s_gam[2 * threadIdx.x + 1]);
            P.theta1_all[(int64_t)(P.it - 1) * P.n_el + iv] = th;
            P.l1[iv] = log(th);
            P.l1c[iv] = log(1.0 - th);
        }
    }
    if (!last) return;
    tl_stamp(P.tl, 2, 6);                      // the block that finished last: its slot 4-6 stamps close the kernel
    const int64_t row = (int64_t)P.it - 1;
    if (threadIdx.x == 32) {
        const double th0 = P.rp_theta0 ? clamp_theta(P.rp_theta0[0])
                                       : beta_from_gammas(s_gam[2 * VPB], s_gam[2 * VPB + 1]);
        P.theta0_all[row] = th0;
        P.scal[SC_THETA0] = th0;
        P.scal[SC_L0] = log(th0);
        P.scal[SC_L0C] = log(1.0 - th0);
    }
    if (threadIdx.x == 64 && P.wise == 0) {
        const double th = P.rp_theta1 ? clamp_theta(P.rp_theta1[0])
                                      : beta_from_gammas(s_gam[2 * VPB + 4], s_gam[2 * VPB + 5]);
        P.theta1_all[row] = th;
        P.l1[0] = log(th);
        P.l1c[0] = log(1.0 - th);
    }
    if (P.tl_hist && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.tl_hist[P.rng_it & 1023u] = t;
    }
    if (threadIdx.x == 96 && P.learn_relax_next && (int64_t)P.it < P.T) {
        const double rr = P.rp_relax_next ? P.rp_relax_next[0]
                                          : beta_from_gammas(s_gam[2 * VPB + 2], s_gam[2 * VPB + 3]);
        P.scal[SC_RELAX] = rr;
        P.scal[SC_ODD1] = log(1.0 - rr) - log(rr);
        P.relax_all[row + 1] = rr;
    }
}

// initial draws (R/clone_id.R:461-463) and state
struct InitParams {
    int64_t N, n_el;
    int K, wise;
    double prior0[2], prior1[2];
    const double* prior1_mat;
    const uint32_t* cfg0;
    uint32_t* mask;
    double* scal; double* l1; double* l1c;
    double* theta0_all; double* theta1_all; double* relax_all;
    uint8_t* config_all; int relax;
    double relax_init;
    int learn_relax_first;     // iteration 2 already learns the relax rate (min_iter tiny)
    double relax_prior[2];
    PhiloxKey key; uint32_t chain;
    const double* rp_theta0; const double* rp_theta1; const double* rp_relax_next;
};

__global__ void k_init(const InitParams P) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P.N) {
        P.mask[i] = P.cfg0[i];
        if (!P.relax)   // Config_new == Config when the relax step is off: trace row 1 holds it too
            for (int k = 0; k < P.K; ++k) P.config_all[i + (int64_t)k * P.N] = (uint8_t)((P.cfg0[i] >> k) & 1u);
    }
    if (i < P.n_el) {
        double th;
        if (P.rp_theta1) th = clamp_theta(P.rp_theta1[i]);
        else {
            const double pa = P.prior1_mat ? P.prior1_mat[i] : P.prior1[0];
            const double pb = P.prior1_mat ? P.prior1_mat[P.n_el + i] : P.prior1[1];
            PhiloxSeq s(P.key, P.chain, 1, STREAM_THETA1, (uint32_t)i);
            th = beta_draw(pa, pb, s);
        }
        P.theta1_all[i] = th;
        P.l1[i] = log(th);
        P.l1c[i] = log(1.0 - th);
    }
    if (i == 0) {
        double th0;
        if (P.rp_theta0) th0 = clamp_theta(P.rp_theta0[0]);
        else {
            PhiloxSeq s(P.key, P.chain, 1, STREAM_THETA0, 0);
            th0 = beta_draw(P.prior0[0], P.prior0[1], s);
        }
        P.theta0_all[0] = th0;
        P.scal[SC_THETA0] = th0;
        P.scal[SC_L0] = log(th0);
        P.scal[SC_L0C] = log(1.0 - th0);
        double rr = P.relax_init;
        if (P.learn_relax_first) {
            if (P.rp_relax_next) rr = P.rp_relax_next[0];
            else {
                PhiloxSeq s(P.key, P.chain, 2, STREAM_RELAX, 0);
                rr = beta_draw(P.relax_prior[0], P.relax_prior[1] + (double)P.N * (double)(P.K - 1), s);
            }
            P.relax_all[1] = rr;
        }
        P.scal[SC_RELAX] = rr;
        P.scal[SC_ODD1] = log(1.0 - rr) - log(rr);
    }
}

template <int KP, int G, typename VIdx>
void launch_cells_t(const CellsParams& P, bool smem_tab, size_t smem, int grid, cudaStream_t st) {
    if (smem_tab) {
        auto kern = k_cells<KP, G, VIdx, true>;
        CDL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, CELLS_THREADS, smem, st>>>(P);
    } else {
        k_cells<KP, G, VIdx, false><<<grid, CELLS_THREADS, 0, st>>>(P);
    }
}

template <int KP, int G>
void launch_cells_v(const Donor& d, const CellsParams& P, bool smem_tab, size_t smem, int grid, cudaStream_t st) {
    if (d.vidx32) launch_cells_t<KP, G, uint32_t>(P, smem_tab, smem, grid, st);
    else launch_cells_t<KP, G, uint16_t>(P, smem_tab, smem, grid, st);
}

template <int KP>
void launch_stats_t(const StatsParams& P, unsigned grid, size_t smem, bool units, cudaStream_t st) {
    auto kern = units ? k_stats_units<KP> : k_stats_stream<KP>;
    CDL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_dep(kern, dim3(grid), dim3(STATS_THREADS), smem, st, P);
}

int pad_k(int K) { return K <= 4 ? 4 : (K <= 8 ? 8 : (K <= 16 ? 16 : 32)); }

}  // namespace

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("CDL_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

int cells_smem_bytes(int64_t N, int K) {
    (void)K;
    const int64_t b = N * (int64_t)(sizeof(double2) + sizeof(uint32_t));
    return b > (int64_t)0x7fffffff ? 0x7fffffff : (int)b;
}

void launch_init_chain(const SweepArgs& a, ChainDev& c, cudaStream_t st) {
    InitParams P{};
    P.N = a.N; P.n_el = a.wise == 1 ? a.N : (a.wise == 2 ? 0 : 1);   // element: k_el_theta<INIT> draws theta1
    P.K = a.K; P.wise = a.wise;
    P.prior0[0] = a.prior0[0]; P.prior0[1] = a.prior0[1];
    P.prior1[0] = a.prior1[0]; P.prior1[1] = a.prior1[1];
    P.prior1_mat = a.prior1_mat; P.cfg0 = a.cfg0; P.mask = c.mask;
    P.scal = c.scal; P.l1 = c.l1; P.l1c = c.l1c;
    P.theta0_all = c.theta0_all; P.theta1_all = c.theta1_all; P.relax_all = c.relax_all;
    P.config_all = c.config_all; P.relax = a.relax;
    P.relax_init = a.relax_fixed_set ? a.relax_fixed : a.relax_prior[0] / (a.relax_prior[0] + a.relax_prior[1]);
    P.learn_relax_first = a.learn_relax_next;
    P.relax_prior[0] = a.relax_prior[0]; P.relax_prior[1] = a.relax_prior[1];
    P.key = a.key; P.chain = a.chain;
    P.rp_theta0 = a.rp_theta0; P.rp_theta1 = a.rp_theta1; P.rp_relax_next = a.rp_relax_next;
    const int64_t n = std::max<int64_t>(a.N, 1);
    k_init<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P);
    CDL_CUDA(cudaGetLastError());
}

void launch_cells(const Donor& d, const SweepArgs& a, ChainDev& c, int sm_count, cudaStream_t st) {
    CellsParams P{};
    P.rowgrp = d.c_rowgrp; P.vidx = d.c_vidx; P.ca = d.c_a; P.cb = d.c_b;
    P.M = d.M; P.N = d.N; P.cell_offset = d.cell_offset; P.K = a.K;
    P.n_tab = a.wise == 1 ? (int)a.N : 1;
    P.l1 = c.l1; P.l1c = c.l1c; P.scal = c.scal; P.mask = c.mask;
    for (int k = 0; k < KMAX; ++k) P.logPsi[k] = a.logPsi[k];
    const int64_t row = (int64_t)a.it - 1;
    P.prob_row = c.prob_all ? c.prob_all + row * d.M * a.K : nullptr;
    P.prob_acc = c.prob_acc;
    P.prob_sq = c.prob_sq;
    P.z = c.z;
    P.assign_row = c.assign_all ? c.assign_all + row * d.M : nullptr;
    P.rp_u = a.rp_u_assign;
    P.key = a.key; P.chain = a.chain; P.it = a.rng_it;
    P.chg_list = a.incremental ? c.chg_list : nullptr; P.chg_old = c.chg_old; P.chg_count = c.chg_ctl;
    const size_t smem = (size_t)cells_smem_bytes(d.N, a.K);
    const bool smem_tab = smem <= 200 * 1024;
    P.tab_in_smem = smem_tab;
    P.zero = 0;
    int ctas_per_sm = 1;
    if (smem_tab) {
        ctas_per_sm = (int)std::min<size_t>(4, (220 * 1024) / std::max<size_t>(smem + 1024, 1));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    } else {
        ctas_per_sm = 4;
        if (!c.tab_g) throw Error(4, "global weight table not allocated");
        k_build_tab<<<(unsigned)((d.N + 255) / 256), 256, 0, st>>>(d.N, P.n_tab, c.l1, c.l1c, c.scal, c.tab_g);
        P.l1 = reinterpret_cast<const double*>(c.tab_g);
    }
    // lanes per cell: the smallest group that still gives every padded clone a lane.  Several cells share
    // a warp, which amortises the per-cell reduction / softmax / draw instructions; a lane then streams
    // rows_groups / G groups of 8 entries with the software prefetch.
    const int KP = pad_k(a.K);
    int G = KP < 8 ? 8 : KP;
    if (const char* e = getenv("CDL_CELLS_G")) {
        const int g = atoi(e);
        if ((g == 8 || g == 16 || g == 32) && g >= KP) G = g;
    }
    const int cpw = 32 / G;
    const int64_t warps_needed = (d.M + cpw - 1) / cpw;
    int64_t grid = (int64_t)sm_count * ctas_per_sm;
    const int64_t max_useful = (warps_needed + CELLS_THREADS / 32 - 1) / (CELLS_THREADS / 32);
    if (grid > max_useful) grid = std::max<int64_t>(max_useful, 1);
    const int code = KP * 100 + G;
    switch (code) {
        case 408: launch_cells_v<4, 8>(d, P, smem_tab, smem, (int)grid, st); break;
        case 416: launch_cells_v<4, 16>(d, P, smem_tab, smem, (int)grid, st); break;
        case 432: launch_cells_v<4, 32>(d, P, smem_tab, smem, (int)grid, st); break;
        case 808: launch_cells_v<8, 8>(d, P, smem_tab, smem, (int)grid, st); break;
        case 816: launch_cells_v<8, 16>(d, P, smem_tab, smem, (int)grid, st); break;
        case 832: launch_cells_v<8, 32>(d, P, smem_tab, smem, (int)grid, st); break;
        case 1616: launch_cells_v<16, 16>(d, P, smem_tab, smem, (int)grid, st); break;
        case 1632: launch_cells_v<16, 32>(d, P, smem_tab, smem, (int)grid, st); break;
        default: launch_cells_v<32, 32>(d, P, smem_tab, smem, (int)grid, st); break;
    }
    CDL_CUDA(cudaGetLastError());
}

// CTA -> cell block plan of k_stats for a grid of `ctas` CTAs (n_cb <= ctas): CTAs in proportion to the blocks'
// entries, at least one each.  Layout: [ctas] home block of every CTA, [n_cb] first CTA, [n_cb] CTA count.
std::vector<int> stats_plan(const std::vector<int64_t>& groups_per_cb, int ctas) {
    const int n_cb = (int)groups_per_cb.size();
    std::vector<int> cnt((size_t)n_cb, 1);
    int64_t total = 0;
    for (int64_t g : groups_per_cb) total += g;
    int left = ctas - n_cb;
    if (total > 0 && left > 0) {
        std::vector<double> want((size_t)n_cb);
        int given = 0;
        for (int c = 0; c < n_cb; ++c) {
            want[(size_t)c] = (double)groups_per_cb[(size_t)c] / (double)total * (double)ctas;
            const int extra = std::max(0, (int)want[(size_t)c] - 1);
            cnt[(size_t)c] += extra; given += extra;
        }
        for (int rest = left - given; rest > 0; --rest) {        // hand out the remainder by largest deficit
            int best = 0; double bd = -1e300;
            for (int c = 0; c < n_cb; ++c) { const double df = want[(size_t)c] - cnt[(size_t)c]; if (df > bd) { bd = df; best = c; } }
            ++cnt[(size_t)best];
        }
    }
    std::vector<int> plan((size_t)ctas + 2 * (size_t)n_cb);
    int b = 0;
    for (int c = 0; c < n_cb; ++c) {
        plan[(size_t)ctas + c] = b;
        plan[(size_t)ctas + n_cb + c] = cnt[(size_t)c];
        for (int q = 0; q < cnt[(size_t)c]; ++q) plan[(size_t)b++] = c;
    }
    return plan;
}

// CDL_CARRY_SUM=0: keep summing the peers' tables every sweep (A/B switch)
static bool carry_sum_enabled() {
    static const bool on = [] { const char* e = getenv("CDL_CARRY_SUM"); return !(e && e[0] == '0'); }();
    return on;
}
static PeerSignal peer_signal(const SweepArgs& a) {
    PeerSignal S{};
    if (a.p2p && a.p2p->n_peer > 0) {
        const P2PState& q = *a.p2p;
        S.n_peer = q.n_peer; S.my_rank = q.my_rank; S.epoch = q.epoch;
        for (int r = 0; r < q.n_peer; ++r) S.peer_flags[r] = q.peer_flags[r];
    }
    return S;
}

void launch_stats(const Donor& d, const SweepArgs& a, ChainDev& c, int sm_count, cudaStream_t st) {
    // c.sab is zero on entry: k_table clears it while it reads it (and the chain starts with a zeroed table)
    StatsParams P{};
    P.sig = peer_signal(a);
    P.done_ticket = c.stats_ticket;
    P.ctl = nullptr;
    if (a.incremental) {
        P.ctl = c.chg_ctl;
        P.threshold = (unsigned int)std::max<int64_t>(1, d.M / 32);      // ~3 % of the cells
        P.zbuf = c.sab_zero;
        P.rowgrp = d.c_rowgrp; P.vidx = d.c_vidx; P.ca = d.c_a; P.cb = d.c_b; P.vidx32 = d.vidx32 ? 1 : 0;
        P.chg_list = c.chg_list; P.chg_old = c.chg_old;
    }
    P.seggrp = d.v_seggrp; P.flush = d.v_flush; P.vc = d.v_cell; P.va = d.v_a; P.vb = d.v_b; P.z = c.z; P.sab = c.sab;
    P.N = d.N; P.M = d.M; P.K = a.K; P.n_cb = d.n_cb;
    P.tl = c.tl;
    const int KP = pad_k(a.K);
    const int ncell_max = (int)std::min<int64_t>(CELL_BLOCK, d.M);
    const size_t smem = sizeof(uint32_t) * STATS_WARPS * KP * 32 + (size_t)((ncell_max + 15) / 16 * 16);
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (220 * 1024) / (smem + 1024)));
    const int ctas = sm_count * ctas_per_sm;
    if (d.n_cb <= ctas) {
        Donor& dm = const_cast<Donor&>(d);          // lazily built launch plan, cached with the donor
        if (!dm.v_plan || dm.v_plan_ctas != ctas) {
            std::vector<uint32_t> ends((size_t)d.n_cb + 1);
            for (int cb = 0; cb <= d.n_cb; ++cb)
                CDL_CUDA(cudaMemcpyAsync(&ends[(size_t)cb], d.v_seggrp + (int64_t)cb * d.N, 4, cudaMemcpyDeviceToHost, st));
            CDL_CUDA(cudaStreamSynchronize(st));
            std::vector<int64_t> gp((size_t)d.n_cb);
            for (int cb = 0; cb < d.n_cb; ++cb) gp[(size_t)cb] = (int64_t)ends[(size_t)cb + 1] - (int64_t)ends[(size_t)cb];
            const std::vector<int> plan = stats_plan(gp, ctas);
            cudaFree(dm.v_plan);
            dm.v_plan = nullptr;
            CDL_CUDA(cudaMalloc(&dm.v_plan, sizeof(int) * plan.size()));
            CDL_CUDA(cudaMemcpyAsync(dm.v_plan, plan.data(), sizeof(int) * plan.size(), cudaMemcpyHostToDevice, st));
            CDL_CUDA(cudaStreamSynchronize(st));
            dm.v_plan_ctas = ctas;
        }
        P.plan = d.v_plan;
    }
    // Many segments per warp (a whole C4): the unit-pulling kernel; few (a 1/8 shard, C3): the streaming kernel
    // with one static range per warp.
    const int64_t warps = (int64_t)ctas * STATS_WARPS;
    const int64_t segs = d.N * (int64_t)d.n_cb;
    bool units = segs >= 16 * warps;
    if (const char* e = getenv("CDL_STATS_KERNEL")) units = e[0] == 'u';
    P.work = c.stats_work;
    P.chunks_per_warp = 1;
    if (units) {
        // work units per segment: at least ~8 units per warp, no unit shorter than ~64 groups
        int64_t parts = (8 * warps + segs - 1) / segs;
        const int64_t avg_groups = std::max<int64_t>(1, d.v_groups / std::max<int64_t>(segs, 1));
        parts = std::min<int64_t>(parts, std::max<int64_t>(1, avg_groups / 64));
        P.chunks_per_warp = (int)std::max<int64_t>(1, std::min<int64_t>(parts, 64));
        // c.stats_work is zero on entry: k_table's last block re-arms it (and the chain starts with it zeroed)
    }
    const unsigned grid = (unsigned)ctas;
    switch (KP) {
        case 4: launch_stats_t<4>(P, grid, smem, units, st); break;
        case 8: launch_stats_t<8>(P, grid, smem, units, st); break;
        case 16: launch_stats_t<16>(P, grid, smem, units, st); break;
        default: launch_stats_t<32>(P, grid, smem, units, st); break;
    }
    CDL_CUDA(cudaGetLastError());
}

void launch_table(const Donor& d, const SweepArgs& a, ChainDev& c, cudaStream_t st) {
    TableParams P{};
    P.N = d.N; P.K = a.K; P.wise = a.wise; P.relax = a.relax; P.keep_base = a.keep_base;
    P.learn_relax_next = a.learn_relax_next;
    for (int q = 0; q < 2; ++q) {
        P.prior0[q] = a.prior0[q]; P.prior1[q] = a.prior1[q]; P.relax_prior[q] = a.relax_prior[q];
    }
    P.prior1_mat = a.prior1_mat;
    P.n_el = a.wise == 1 ? d.N : (a.wise == 2 ? d.nnz : 1);
    P.cfg0 = a.cfg0; P.mask = c.mask; P.sab = c.sab; P.scal = c.scal; P.l1 = c.l1; P.l1c = c.l1c;
    P.sd1 = a.wise == 2 ? c.sd1 : nullptr;
    P.W_log = a.W_log;
    P.theta0_all = c.theta0_all; P.theta1_all = c.theta1_all; P.loglik_all = c.loglik_all;
    P.relax_all = c.relax_all; P.config_all = c.config_all;
    P.T = a.T; P.it = a.it; P.rng_it = a.rng_it; P.key = a.key; P.chain = a.chain;
    P.rp_u_config = a.rp_u_config; P.rp_theta0 = a.rp_theta0; P.rp_theta1 = a.rp_theta1;
    P.rp_relax_next = a.rp_relax_next;
    P.part_d = c.part_d; P.part_u = c.part_u; P.ticket = c.ticket;
    P.sell_work = c.work;
    P.stats_work = c.stats_work; P.n_cb = (int)d.n_cb;
    P.sab_next = (a.incremental || a.wise == 2) ? nullptr : c.sab;
    P.carry_local = 0;
    P.dbg_s34 = c.dbg_s34; P.dbg_tot = c.dbg_s34 ? c.dbg_s34 + 2 * d.N : nullptr;
    P.n_peer = 0;
    P.tl = c.tl; P.tl_hist = c.tl_hist;
    if (a.p2p && a.p2p->n_peer > 0) {
        const P2PState& q = *a.p2p;
        P.n_peer = q.n_peer;
        for (int r = 0; r < q.n_peer; ++r) {
            P.peer_sab[r] = q.peer_arena[r] + (size_t)(q.epoch & 1ull) * q.table_elems;
            P.peer_rank[r] = q.peer_rank[r];
        }
        P.sab_next = q.arena + (size_t)((q.epoch + 1ull) & 1ull) * q.table_elems;   // not read by any peer any more
        P.carry_local = a.incremental ? 1 : 0;
        P.my_flags = q.my_flags;
        P.epoch = q.epoch;
        P.p2p_error = q.error;
        P.g_own = (a.incremental && carry_sum_enabled()) ? q.arena + 2 * q.table_elems : nullptr;
        P.ctl = c.chg_ctl;
    }
    const int KP = pad_k(a.K);
    const int vpb = TABLE_THREADS / KP;
    const unsigned grid = (unsigned)((d.N + vpb - 1) / vpb);
    switch (KP) {
        case 4: launch_dep(k_table<4>, dim3(grid), dim3(TABLE_THREADS), 0, st, P); break;
        case 8: launch_dep(k_table<8>, dim3(grid), dim3(TABLE_THREADS), 0, st, P); break;
        case 16: launch_dep(k_table<16>, dim3(grid), dim3(TABLE_THREADS), 0, st, P); break;
        default: launch_dep(k_table<32>, dim3(grid), dim3(TABLE_THREADS), 0, st, P); break;
    }
    CDL_CUDA(cudaGetLastError());
}

}  // namespace cdl
