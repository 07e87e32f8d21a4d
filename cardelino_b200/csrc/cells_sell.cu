// Slice-major cell kernel for large donors (the BASELINE C3/C4 shapes): logLik thin SpMM -> softmax ->
// categorical draw (R/clone_id.R:470-497) with ONE LANE PER CELL.
//
// Layout (built once per donor from the row layout by donor_build_sell): cells are sorted by row length
// inside windows of SELL_SIGMA cells and cut into slices of 32; slice s owns group slots
// [slice_off[s], slice_off[s+1]) and slot g stores, for each of the 32 cells, one 8-entry group
// (variant index, A, B as uint16) at [(g * 32 + lane)] -- so one warp-wide 128-bit load per array reads
// 512 contiguous bytes, each lane walks its own cell, and nothing is reduced across lanes: the per-cell
// softmax / draw epilogue costs a few warp instructions per cell instead of hundreds.
//
// The per-variant weights (log th1_i - log th0, log(1-th1_i) - log(1-th0)) sit in shared memory as one
// double2 per variant.  For K <= 16 the Config_new row of the variant rides in the lowest mantissa byte
// of each of the two doubles (clones 0-7 in the first, 8-15 in the second): one 128-bit shared load per
// entry instead of two loads, and R2P reads the predicate bits straight from those registers.  The
// perturbation is <= 2^-44 relative per weight (~1e-10 absolute on a logit of 500 entries), four orders
// below the 1e-6 parity bar; the Config flips, theta draws and the logLik trace (k_table) use the exact
// weights.
#include "internal.h"
#include "cells.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <queue>
#include <vector>

namespace cdl {

namespace {

// launch shapes: (threads per CTA, group slots per lane loaded ahead of the ones being accumulated)
constexpr int SELL_SIGMA = 4096;   // sorting window in cells (multiple of 32)

struct SellParams {
    CellsParams c;
    const int32_t* assign;       // one-round shards: slice of every (block, warp) [grid * warps], -1 = none; else nullptr
    const uint32_t* slice_off;   // [n_slices + 1] group-slot offsets
    const uint32_t* perm;        // [n_slices * 32] local cell id of each lane, 0xffffffff = empty lane
    const uint4* v8;
    const uint4* a8;
    const uint4* b8;
    int64_t n_slices;
    unsigned int* work;          // counter of the work items after every warp's static first one; zero on entry
                                 // (k_table's last block re-arms it)
    // small shards: a slice's groups are split over `parts` warps (work item = slice * parts + part); each
    // part leaves its partial logits in `scratch`, the part that finishes last (ticket) adds them up in part
    // order -- deterministic -- and runs the epilogue
    int hi_shift;                // the odd entries of a group store their variant id << hi_shift (Donor::s_hi_shift)
    int parts;
    double* scratch;             // [n_slices][parts][32][KP]
    unsigned int* tickets;       // [n_slices], zero between launches
};

// acc[k] += bit_q(bits) ? del : 0 for NK clones starting at K0.  Written as a branch around the add in
// volatile PTX: ptxas if-converts that into a PREDICATED DADD (@P DADD) with the predicates produced seven at
// a time by R2P -- about 1.2 instructions per clone.  (A predicated add written directly, or a ?: in C, is
// turned into two FSELs + an unconditional DADD, and the high-word-select form needs a SEL and a register
// move per clone.)
template <int K0, int NK, int KP>
__device__ __forceinline__ void masked_add8(double (&acc)[KP], double del, uint32_t bits) {
#pragma unroll
    for (int q = 0; q < NK; ++q) {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                     "and.b32 t, %2, %3;\n\tsetp.eq.u32 p, t, 0;\n\t@p bra SKIP%=;\n\t"
                     "add.f64 %0, %0, %1;\n\tSKIP%=:\n\t}"
                     : "+d"(acc[K0 + q])
                     : "d"(del), "r"(bits), "r"(1u << q));
    }
}

// ---- the fast path: weight table in shared memory, K <= 16, uint16 variant ids (every BASELINE shape) ----
// The inner loop is bound by instruction issue and the fp64 pipe together (15-16 predicated DADDs + DMUL + DFMA per
// covered entry), so everything else per entry is cut to the bone (profiles/r2_sass_k_cells_sell_inner_loop.txt):
//   * the odd entries of a group store their variant id PRE-SHIFTED by two (Donor::s_hi_shift): the table address of
//     the high half is ONE LEA.HI (base + (w >> 14)); the low half costs PRMT + LEA.HI;
//   * counts are converted straight from the register halves (I2F.F64.U16 .H0 / .H1);
//   * the Config bits ride in the weights' low mantissa bytes in R2P order: clones 1-7 (0-6 without the base-clone
//     shortcut) in bits 0-6 of du's byte, the next seven in bits 0-6 of dv's byte, the one or two left over in
//     bit 7 of each -- two R2P + one LOP3 per entry, no shifts;
//   * two groups per loop trip with ping-pong registers: no register rotation, 32-bit loop arithmetic.
template <int KP, bool SKIP0>
__host__ __device__ constexpr int sell_bitpos(int k) {
    const int a = k - (SKIP0 ? 1 : 0);          // rank among the clones that get a masked add
    return a < 0 ? 31 : (a < 7 ? a : (a < 14 ? 8 + (a - 7) : (a == 14 ? 7 : 15)));
}
// Config row (clone bit mask) -> the two mask bytes in sell_bitpos order
template <int KP, bool SKIP0>
__device__ __forceinline__ uint32_t sell_pack_mask(uint32_t m) {
    uint32_t o = 0;
#pragma unroll
    for (int k = SKIP0 ? 1 : 0; k < KP; ++k) o |= ((m >> k) & 1u) << sell_bitpos<KP, SKIP0>(k);
    return o;
}
__device__ __forceinline__ double u16lo_to_double(uint32_t w) {
    double r;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.rn.f64.u16 %0, l;\n\t}" : "=d"(r) : "r"(w));
    return r;
}
__device__ __forceinline__ double u16hi_to_double(uint32_t w) {
    double r;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.rn.f64.u16 %0, h;\n\t}" : "=d"(r) : "r"(w));
    return r;
}
template <int KP, bool SKIP0>
__device__ __forceinline__ void sell_group_fast(double (&acc)[KP], const Grp<uint16_t>& q, uint32_t tab_sa) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t vw = e < 2 ? q.v.x : (e < 4 ? q.v.y : (e < 6 ? q.v.z : q.v.w));
        const uint32_t aw = e < 2 ? q.a.x : (e < 4 ? q.a.y : (e < 6 ? q.a.z : q.a.w));
        const uint32_t bw = e < 2 ? q.b.x : (e < 4 ? q.b.y : (e < 6 ? q.b.z : q.b.w));
        // byte offset of the variant's double2: high half holds 4 * id (bits 14-15 of the low half are zero, ids
        // are below 2^14), low half the plain id moved to the top of a word first
        const uint32_t off = (e & 1) ? (vw >> 14) : (__byte_perm(vw, 0u, 0x1044) >> 12);
        double du, dv;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(du), "=d"(dv) : "r"(tab_sa + off));
        const double da = (e & 1) ? u16hi_to_double(aw) : u16lo_to_double(aw);
        const double db = (e & 1) ? u16hi_to_double(bw) : u16lo_to_double(bw);
        const double del = fma(da, du, db * dv);
        // both mask bytes in ONE integer register: ptxas then produces the predicates with R2P (seven per
        // instruction); testing the low words of du / dv in place makes it emit one LOP3 per clone instead
        const uint32_t m = __byte_perm((uint32_t)__double2loint(du), (uint32_t)__double2loint(dv), 0x7740);
#pragma unroll
        for (int k = SKIP0 ? 1 : 0; k < KP; ++k) {
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                         "and.b32 t, %2, %3;\n\tsetp.eq.u32 p, t, 0;\n\t@p bra SKIP%=;\n\t"
                         "add.f64 %0, %0, %1;\n\tSKIP%=:\n\t}"
                         : "+d"(acc[k])
                         : "d"(del), "r"(m), "r"(1u << sell_bitpos<KP, SKIP0>(k)));
        }
    }
}

template <int KP, typename VIdx, bool SKIP0>
__device__ __forceinline__ void sell_group_fast_any(double (&acc)[KP], const Grp<VIdx>& q, uint32_t tab_sa) {
    if constexpr (sizeof(VIdx) == 2) sell_group_fast<KP, SKIP0>(acc, q, tab_sa);
}

template <int KP, typename VIdx, bool SMEM, bool PACKED, bool SKIP0>
__device__ __forceinline__ void sell_group(double (&acc)[KP], const Grp<VIdx>& q, uint32_t tab_sa, uint32_t msk_sa,
                                           const double2* __restrict__ tab_g, const uint32_t* __restrict__ msk_g,
                                           int hi_shift) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t aw = e < 2 ? q.a.x : (e < 4 ? q.a.y : (e < 6 ? q.a.z : q.a.w));
        const uint32_t bw = e < 2 ? q.b.x : (e < 4 ? q.b.y : (e < 6 ? q.b.z : q.b.w));
        const uint32_t a = (e & 1) ? (aw >> 16) : (aw & 0xffffu);
        const uint32_t b = (e & 1) ? (bw >> 16) : (bw & 0xffffu);
        double du, dv;
        if (PACKED) {
            const uint32_t i = (e & 1) ? (q.idx(e) >> hi_shift) : q.idx(e);
            if (SMEM) {
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(du), "=d"(dv) : "r"(tab_sa + (i << 4)));
            } else {
                const double2 t = __ldg(tab_g + i);
                du = t.x; dv = t.y;
            }
            // counts -> double with I2F.F64.U16 (converts a register half directly, on the conversion pipe):
            // the 2^52 trick of the row kernel would put two more DADDs per entry on the fp64 pipe, which is
            // what bounds this kernel
            const double del = fma((double)a, du, (double)b * dv);
            // gather the two mask bytes into ONE integer register first: ptxas then turns the bit tests into
            // R2P (seven predicates per instruction); testing the low words of du / dv in place makes it
            // emit one LOP3 per clone instead
            const uint32_t m = __byte_perm((uint32_t)__double2loint(du), (uint32_t)__double2loint(dv), 0x7740);
            // SKIP0: the base clone's Config column is zero and cannot flip (keep_base_clone), so its logit gets
            // no contribution at all -- one masked add fewer per entry
            if (SKIP0) masked_add8<1, (KP < 8 ? KP : 8) - 1, KP>(acc, del, m >> 1);
            else masked_add8<0, (KP < 8 ? KP : 8), KP>(acc, del, m);
            if (KP > 8) masked_add8<8, KP - 8, KP>(acc, del, m >> 8);
        } else {
            uint32_t m;
            load_tab<SMEM>(tab_sa, msk_sa, tab_g, msk_g, (e & 1) ? (q.idx(e) >> hi_shift) : q.idx(e), du, dv, m);
            const double del = fma(u16_to_double(a), du, u16_to_double(b) * dv);
            masked_add8<0, 8, KP>(acc, del, m);
            masked_add8<8, 8, KP>(acc, del, m >> 8);
            if (KP > 16) {
                masked_add8<16, 8, KP>(acc, del, m >> 16);
                masked_add8<24, 8, KP>(acc, del, m >> 24);
            }
        }
    }
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only) and the wait for all of this thread's copies
__device__ __forceinline__ void cp_async16(uint32_t dst_sa, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_sa), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t sa) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(sa));
    return r;
}
constexpr int SELL_STAGE_BYTES = 3 * 32 * 16;      // one group slot of a warp: variant ids, A, B (16 bytes per lane each)

// ASYNC (fast path only): the next group slot of a warp travels HBM -> shared memory with cp.async while the current one
// is accumulated from registers.  The copy is issued where it is written -- a whole group (~3000 cycles of this warp's
// share of the fp64 pipe) before its data is needed -- which ptxas does not do for plain loads at this register budget
// (it sinks them to a few instructions before their use: 19 % of the stall samples on that first use), and the look-ahead
// group occupies shared memory instead of 12 registers.
template <int KP, typename VIdx, bool SMEM, int SELL_THREADS, int SELL_DEPTH, bool SPLIT, bool SKIP0, bool ASYNC = false>
__global__ void __launch_bounds__(SELL_THREADS, 1) k_cells_sell(const SellParams P) {
    constexpr bool PACKED = KP <= 16;
    constexpr bool FAST = SMEM && PACKED && sizeof(VIdx) == 2;     // see sell_group_fast
    static_assert(!ASYNC || (FAST && !SPLIT), "the cp.async pipeline is written for the fast path");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_trigger();
    const CellsParams& C = P.c;
    tl_stamp(C.tl, 0, 0);
    if (C.tl && threadIdx.x == 0 && blockIdx.x < TL_ROWS) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        C.tl[(size_t)blockIdx.x * TL_SLOTS + 4] = smid;
    }
    const int lane = threadIdx.x & 31;
    // Work items (slices, or 1/parts of one) are dealt dynamically, except each warp's FIRST one, which is its
    // own index: its header and first group slots are donor layout only, so under a programmatic dependent
    // launch they are fetched here, while the table kernel is still running.
    const unsigned int n_static = gridDim.x * (blockDim.x >> 5);
    // (strided by the grid, not blocked by CTA: slices are sorted by length inside 4096-cell windows, and on a shard
    // that is done in one round a CTA holding 28 neighbours would be all-long or all-short)
    unsigned int s_u = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
    // A shard that is done in ONE round (at most one slice per warp: 27 per SM on a 1/8 shard of C4) has no dynamic part
    // to even things out, and the four sub-partitions of an SM each share their fp64 pipe among their own warps: the
    // host has dealt the slices to the (SM, sub-partition) bins so that every bin carries the same number of group
    // slots (launch_cells_sell: longest-first into the least loaded bin), this warp's slice is assign[block][warp].
    if (!SPLIT && P.assign) {
        const int32_t sl = P.assign[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)];
        s_u = sl < 0 ? 0xffffffffu : (unsigned int)sl;
    }
    int64_t s = 0, g0 = 0, g1 = 0;
    int part = 0;
    uint32_t jl = 0xffffffffu;
    Grp<VIdx> q[SELL_DEPTH];
    // this lane's 16-byte cells of its warp's staging slot (behind the weight table): variant ids, A at + 512, B at + 1024
    const uint32_t stage_sa = ASYNC ? (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)(sizeof(double2) * (size_t)C.N) +
                                          (uint32_t)((threadIdx.x >> 5) * SELL_STAGE_BYTES + lane * 16)
                                    : 0u;
    // `dep` is zero, but derived from what the previous occupant of the slot was read into: the copy cannot be issued
    // before those reads have returned
    auto stage_issue = [&](uint32_t g, uint32_t dep) {
        const size_t i = (size_t)g * 32 + (size_t)lane;
        cp_async16(stage_sa + dep, reinterpret_cast<const uint4*>(C.vidx) + i);
        cp_async16(stage_sa + 512u + dep, P.a8 + i);
        cp_async16(stage_sa + 1024u + dep, P.b8 + i);
    };
    // header of item s_u and its first SELL_DEPTH group slots (the look-ahead index is clamped to the slice's last
    // slot so the loads need no predicate: a few redundant loads per slice); false when the items are used up
    auto item_begin = [&]() -> bool {
        const int parts = SPLIT ? P.parts : 1;
        s = (int64_t)(s_u / (unsigned int)parts);
        part = (int)(s_u - (unsigned int)s * (unsigned int)parts);
        if (s >= P.n_slices) return false;
        g0 = P.slice_off[s]; g1 = P.slice_off[s + 1];
        if (SPLIT && parts > 1) {
            const int64_t len = g1 - g0;
            g1 = g0 + (len * (part + 1)) / parts;
            g0 = g0 + (len * part) / parts;
        }
        jl = P.perm[s * 32 + lane];
        if (g0 < g1) {
            if constexpr (ASYNC) {
                stage_issue((uint32_t)g0, 0u);
            } else {
#pragma unroll
                for (int h = 0; h < SELL_DEPTH; ++h) q[h].load(C.vidx, P.a8, P.b8, min(g0 + h, g1 - 1) * 32 + lane);
            }
        }
        return true;
    };
    bool have = item_begin();
    pdl_wait();
    tl_stamp(C.tl, 0, 1);
    uint32_t tab_sa = 0, msk_sa = 0;
    const double2* tab_g = nullptr;
    const uint32_t* msk_g = C.mask;
    if (SMEM) {
        const double l0 = C.scal[SC_L0], l0c = C.scal[SC_L0C];
        double2* s_tab = reinterpret_cast<double2*>(smem_raw);
        uint32_t* s_msk = reinterpret_cast<uint32_t*>(smem_raw + sizeof(double2) * (size_t)C.N);
        // Nobody computes until the table is staged (3-4 us of a 0.11 ms launch on a 1/8 shard of C4 when every trip waits
        // for its own three loads): four trips' loads are issued together, so the CTA pays ~3 L2 round trips, not 11.
        constexpr int STAGE_U = 4;
        for (int64_t i0 = threadIdx.x; i0 < C.N; i0 += (int64_t)STAGE_U * blockDim.x) {
            double w1[STAGE_U], w1c[STAGE_U];
            uint32_t wm[STAGE_U];
#pragma unroll
            for (int u = 0; u < STAGE_U; ++u) {
                const int64_t i = i0 + (int64_t)u * blockDim.x;
                if (i < C.N) {
                    const int64_t t = C.n_tab == 1 ? 0 : i;
                    w1[u] = C.l1[t]; w1c[u] = C.l1c[t]; wm[u] = C.mask[i];
                }
            }
#pragma unroll
            for (int u = 0; u < STAGE_U; ++u) {
                const int64_t i = i0 + (int64_t)u * blockDim.x;
                if (i >= C.N) break;
                double du = w1[u] - l0, dv = w1c[u] - l0c;
                uint32_t m = wm[u];
                if (FAST) m = sell_pack_mask<KP, SKIP0>(m);
                if (PACKED) {
                    du = __hiloint2double(__double2hiint(du), (__double2loint(du) & ~0xff) | (int)(m & 0xffu));
                    dv = __hiloint2double(__double2hiint(dv), (__double2loint(dv) & ~0xff) | (int)((m >> 8) & 0xffu));
                } else {
                    s_msk[i] = m;
                }
                s_tab[i] = make_double2(du, dv);
            }
        }
        __syncthreads();
        tl_stamp(C.tl, 0, 2);
        tab_sa = (uint32_t)__cvta_generic_to_shared(s_tab);
        msk_sa = (uint32_t)__cvta_generic_to_shared(s_msk);
    } else {
        tab_g = reinterpret_cast<const double2*>(C.l1);   // pre-built (packed when K <= 16) table in global memory
    }
    const int K = C.K;
    auto advance = [&]() {
        if (!SPLIT && P.assign) { have = false; return; }          // one round: nothing left to pull
        unsigned int u = 0;
        if (lane == 0) u = atomicAdd(P.work, 1u);
        s_u = __shfl_sync(0xffffffffu, u, 0) + n_static;
        have = item_begin();
    };
    for (; have; advance()) {
        const int parts = SPLIT ? P.parts : 1;
        const bool valid = jl != 0xffffffffu;
        double acc[KP];
#pragma unroll
        for (int q_ = 0; q_ < KP; ++q_) acc[q_] = 0.0;
        if constexpr (ASYNC) {
            const uint32_t ge = (uint32_t)g1, gl = (uint32_t)g1 - 1u;
#pragma unroll 2
            for (uint32_t g = (uint32_t)g0; g < ge; ++g) {
                cp_async_wait_all();
                Grp<VIdx> cur;
                cur.v = lds128(stage_sa); cur.a = lds128(stage_sa + 512u); cur.b = lds128(stage_sa + 1024u);
                // (clamped to the slice's last slot: no predicate, one redundant copy per slice)
                stage_issue(min(g + 1u, gl), cur.b.w & (uint32_t)C.zero);
                sell_group_fast_any<KP, VIdx, SKIP0>(acc, cur, tab_sa);
            }
            cp_async_wait_all();               // the redundant last copy must have landed before the slot is reused
        } else if (FAST) {
            // A ring of SELL_DEPTH + 1 group buffers, SELL_DEPTH + 1 groups per trip so that every index is static (no
            // register rotation): the loads of a group are written SELL_DEPTH groups before it is accumulated.  The
            // look-ahead index is clamped to the slice's last slot, so the loads need no predicate.  (ptxas places
            // the loads where its register budget allows -- at 72 registers one of the two groups is fetched just
            // before its use, which shows as long-scoreboard samples on its first consumer; pinning the loads early
            // with ld.volatile, 24 / 20 / 16 warps with two or three groups of look-ahead, all measured slower:
            // profiles/README.md, "tried and not kept".)
            const uint32_t ge = (uint32_t)g1, gl = (uint32_t)g1 - 1u;
            Grp<VIdx> n;
            for (uint32_t g = (uint32_t)g0; g < ge; g += SELL_DEPTH + 1) {
#pragma unroll
                for (int st = 0; st <= SELL_DEPTH; ++st) {
                    Grp<VIdx>& dst = ((st + SELL_DEPTH) % (SELL_DEPTH + 1)) == SELL_DEPTH ? n : q[(st + SELL_DEPTH) % (SELL_DEPTH + 1)];
                    dst.load(C.vidx, P.a8, P.b8, (int64_t)(min(g + (uint32_t)(st + SELL_DEPTH), gl) * 32u + (uint32_t)lane));
                    if (st == 0 || g + (uint32_t)st < ge)
                        sell_group_fast_any<KP, VIdx, SKIP0>(acc, st == SELL_DEPTH ? n : q[st], tab_sa);
                }
            }
        } else {
        for (int64_t g = g0; g < g1; g += SELL_DEPTH) {
            Grp<VIdx> n[SELL_DEPTH];
#pragma unroll
            for (int h = 0; h < SELL_DEPTH; ++h)
                n[h].load(C.vidx, P.a8, P.b8, min(g + SELL_DEPTH + h, g1 - 1) * 32 + lane);
#pragma unroll
            for (int h = 0; h < SELL_DEPTH; ++h)
                if (h == 0 || g + h < g1) {
                    if (FAST) sell_group_fast_any<KP, VIdx, SKIP0>(acc, q[h], tab_sa);
                    else sell_group<KP, VIdx, SMEM, PACKED, SKIP0>(acc, q[h], tab_sa, msk_sa, tab_g, msk_g, P.hi_shift);
                }
#pragma unroll
            for (int h = 0; h < SELL_DEPTH; ++h) q[h] = n[h];
        }
        }
        if (SPLIT && parts > 1) {
            double* mine = P.scratch + (((size_t)s * parts + part) * 32 + lane) * KP;
#pragma unroll
            for (int k = 0; k < KP; k += 2) __stcg(reinterpret_cast<double2*>(mine + k), make_double2(acc[k], acc[k + 1]));
            __threadfence();
            unsigned int tk = 0;
            if (lane == 0) tk = atomicAdd(P.tickets + s, 1u);
            tk = __shfl_sync(0xffffffffu, tk, 0);
            if (tk != (unsigned int)(parts - 1)) continue;        // another part finishes this slice
            __threadfence();
            if (lane == 0) P.tickets[s] = 0;
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = 0.0;
            for (int q = 0; q < parts; ++q) {                      // fixed order: bit-reproducible
                const double* src = P.scratch + (((size_t)s * parts + q) * 32 + lane) * KP;
#pragma unroll
                for (int k = 0; k < KP; k += 2) {
                    const double2 t = __ldcg(reinterpret_cast<const double2*>(src + k));
                    acc[k] += t.x; acc[k + 1] += t.y;
                }
            }
        }
        // ---- softmax over the lane's own K logits (:479-481) ----
        double mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) { acc[k] += C.logPsi[k]; mx = fmax(mx, acc[k]); }
        double sm = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k < K) { acc[k] = exp(acc[k] - mx); sm += acc[k]; }
            else acc[k] = 0.0;
        }
        const double inv_sm = 1.0 / sm;            // one division per cell; each probability is within an ulp of e / sm
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) acc[k] = acc[k] * inv_sm;
        if (valid) {
            if (C.prob_row) {
                double* o = C.prob_row + (int64_t)jl * K;
#pragma unroll
                for (int k = 0; k < KP; ++k) if (k < K) o[k] = acc[k];
            }
            if (C.prob_acc) {
                double* o = C.prob_acc + (int64_t)jl * K;
#pragma unroll
                for (int k = 0; k < KP; ++k) if (k < K) o[k] += acc[k];
            }
            if (C.prob_sq) {
                double* o = C.prob_sq + (int64_t)jl * K;
#pragma unroll
                for (int k = 0; k < KP; ++k) if (k < K) o[k] += acc[k] * acc[k];
            }
        }
        // ---- sample(seq_len(K), 1, prob = pr): walk the DESCENDING-sorted CDF with one uniform (:493) ----
        double u = 0.5;
        if (valid) u = C.rp_u ? C.rp_u[jl] : philox_uniform(C.key, C.chain, C.it, STREAM_ASSIGN,
                                                            (uint32_t)(C.cell_offset + jl));
        uint32_t used = 0;
        double cum = 0.0, selv = 0.0;
        int sel = 0;
        bool pending = valid;
        for (int r = 0; r < K; ++r) {
            if (!__any_sync(0xffffffffu, pending)) break;
            double bv = -1.0;
            int bk = 0;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const bool take = (k < K) && !((used >> k) & 1u) && (acc[k] > bv);   // ties -> lowest index
                bv = take ? acc[k] : bv;
                bk = take ? k : bk;
            }
            if (pending) {
                used |= 1u << bk;
                cum += bv;
                if (u <= cum || r == K - 1) { sel = bk; selv = bv; pending = false; }
            }
        }
        // the order inside a group of exactly tied probabilities is revsort's; it only matters when the
        // walk stops inside such a group (e.g. cells without any covered variant: all K tie)
        int ntie = 0;
#pragma unroll
        for (int k = 0; k < KP; ++k) ntie += (k < K && acc[k] == selv) ? 1 : 0;
        if (valid && ntie > 1 && selv > 0.0) {
            double pa[KP];
            int perm[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) { pa[k] = acc[k]; perm[k] = k; }
            double tot = 0.0;
            for (int l = 0; l < K; ++l) tot += pa[l];
            for (int l = 0; l < K; ++l) pa[l] /= tot;
            revsort_dev(pa, perm, K);
            for (int l = 1; l < K; ++l) pa[l] += pa[l - 1];
            int jj = 0;
            while (jj < K - 1 && !(u <= pa[jj])) ++jj;
            sel = perm[jj];
        }
        store_label(C, valid, (int64_t)jl, sel);
    }
    if (C.tl && (threadIdx.x & 31) == 0 && blockIdx.x < TL_ROWS) {      // when the block's last warp ran out of slices
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(C.tl + (size_t)blockIdx.x * TL_SLOTS + 3, t);
    }
}

// (du, dv) table in global memory for donors whose table does not fit shared memory; packed like the
// shared-memory one when K <= 16
__global__ void k_build_tab_sell(int64_t N, int n_tab, int packed, const double* __restrict__ l1,
                                 const double* __restrict__ l1c, const double* __restrict__ scal,
                                 const uint32_t* __restrict__ mask, double2* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t t = n_tab == 1 ? 0 : i;
    double du = l1[t] - scal[SC_L0], dv = l1c[t] - scal[SC_L0C];
    if (packed) {
        const uint32_t m = mask[i];
        du = __hiloint2double(__double2hiint(du), (__double2loint(du) & ~0xff) | (int)(m & 0xffu));
        dv = __hiloint2double(__double2hiint(dv), (__double2loint(dv) & ~0xff) | (int)((m >> 8) & 0xffu));
    }
    out[i] = make_double2(du, dv);
}

// ---- layout construction --------------------------------------------------------------------------
__global__ void k_sell_keys(const uint32_t* __restrict__ rowgrp, int64_t M, int64_t slots,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= slots) return;
    const uint32_t len = j < M ? rowgrp[j + 1] - rowgrp[j] : 0u;
    keys[j] = ((uint64_t)(j / SELL_SIGMA) << 32) | (uint64_t)(0xffffffffu - len);   // window asc, length desc
    vals[j] = j < M ? (uint32_t)j : 0xffffffffu;
}

__global__ void k_sell_slice_len(const uint32_t* __restrict__ rowgrp, const uint32_t* __restrict__ perm,
                                 int64_t n_slices, uint32_t* __restrict__ len) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (s > n_slices) return;
    uint32_t l = 0;
    if (s < n_slices) {
        const uint32_t j = perm[s * 32 + lane];
        if (j != 0xffffffffu) l = rowgrp[j + 1] - rowgrp[j];
    }
    l = __reduce_max_sync(0xffffffffu, l);
    if (lane == 0) len[s] = l;      // len[n_slices] = 0 closes the scan
}

// One thread per lane slot.  Inside a cell the order of the entries is free (the logit is a sum), so the row
// is dealt out such that entry slot e of every group of lane l holds a variant with
// (variant mod 8) == (l + e) mod 8: the eight lanes of a quarter-warp then gather eight DISTINCT 16-byte bank
// groups of the shared-memory weight table with each 128-bit load (conflict-free) instead of eight random ones
// (about 2.4-way conflicts).  A class with more entries than the row has groups spills into free slots of
// other classes (a few percent of the entries); empty slots are padded with a zero-count entry of the
// matching class.
template <typename VIdx>
__global__ void k_sell_fill(const uint32_t* __restrict__ rowgrp, const uint32_t* __restrict__ perm,
                            const uint32_t* __restrict__ slice_off, int64_t n_slices, int64_t N,
                            const VIdx* __restrict__ cv, const uint16_t* __restrict__ ca,
                            const uint16_t* __restrict__ cb, VIdx* __restrict__ ov, uint16_t* __restrict__ oa,
                            uint16_t* __restrict__ ob, int hi_shift) {
    const int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (slot >= n_slices * 32) return;
    const int64_t s = slot >> 5;
    const int lane = (int)(slot & 31);
    const int64_t o0 = slice_off[s];
    const int ngs = (int)(slice_off[s + 1] - o0);          // groups of the slice (>= groups of this row)
    const uint32_t j = perm[slot];
    int64_t e0 = 0, e1 = 0;
    if (j != 0xffffffffu) { e0 = (int64_t)rowgrp[j] * GROUP; e1 = (int64_t)rowgrp[j + 1] * GROUP; }
    auto out_index = [&](int g, int e) { return ((o0 + g) * 32 + lane) * GROUP + e; };
    int cursor[GROUP];
#pragma unroll
    for (int e = 0; e < GROUP; ++e) cursor[e] = 0;
    // pass 1: entries whose class column still has room
    for (int64_t q = e0; q < e1; ++q) {
        const uint32_t a = ca[q], b = cb[q];
        if ((a | b) == 0) continue;                        // padding of the row layout
        const uint32_t v = (uint32_t)cv[q];
        const int e = (int)((v - (uint32_t)lane) & 7u);
        if (cursor[e] < ngs) {
            const int64_t o = out_index(cursor[e]++, e);
            ov[o] = (VIdx)(v << ((e & 1) ? hi_shift : 0)); oa[o] = (uint16_t)a; ob[o] = (uint16_t)b;
        }
    }
    // pass 2: the overflow of over-full classes goes to the free slots of the others
    int placed[GROUP], fill[GROUP];
#pragma unroll
    for (int e = 0; e < GROUP; ++e) { placed[e] = 0; fill[e] = cursor[e]; }
    int fe = 0;
    for (int64_t q = e0; q < e1; ++q) {
        const uint32_t a = ca[q], b = cb[q];
        if ((a | b) == 0) continue;
        const uint32_t v = (uint32_t)cv[q];
        const int e = (int)((v - (uint32_t)lane) & 7u);
        if (placed[e]++ < ngs) continue;                   // went out in pass 1
        while (fe < GROUP && fill[fe] >= ngs) ++fe;
        if (fe >= GROUP) break;                            // cannot happen: the row has <= 8 * ngs entries
        const int64_t o = out_index(fill[fe]++, fe);
        ov[o] = (VIdx)(v << ((fe & 1) ? hi_shift : 0)); oa[o] = (uint16_t)a; ob[o] = (uint16_t)b;
    }
    // padding: zero counts, a variant id of the slot's class so that it gathers conflict-free too
    for (int e = 0; e < GROUP; ++e) {
        const uint32_t want = (uint32_t)(lane + e) & 7u;
        const uint32_t vpad = (int64_t)want < N ? want : 0u;
        for (int g = fill[e]; g < ngs; ++g) {
            const int64_t o = out_index(g, e);
            ov[o] = (VIdx)(vpad << ((e & 1) ? hi_shift : 0)); oa[o] = 0; ob[o] = 0;
        }
    }
}

// Warps per CTA (one CTA per SM): slices are handed out dynamically, one per warp at a time, so a launch takes
// ceil(slices per SM / warps) rounds.  With few slices per SM (a 1/8 shard of C4 has 26) a full CTA of 24
// warps would run one full round and a nearly empty second one; using just enough warps to fill the same
// number of rounds evenly removes that tail.
inline int sell_warps(int64_t n_items, int grid, int max_warps) {
    const int64_t per_sm = (n_items + grid - 1) / grid;
    const int64_t rounds = (per_sm + max_warps - 1) / max_warps;
    int64_t w = (per_sm + rounds - 1) / rounds;
    if (w < 4) w = 4;
    if (w > max_warps) w = max_warps;
    return (int)w;
}

template <int KP, typename VIdx, int THREADS, int DEPTH, bool SKIP0>
void launch_sell_t(const SellParams& P, bool smem_tab, size_t smem, int grid, cudaStream_t st) {
    const int threads = 32 * sell_warps(P.n_slices * P.parts, grid, THREADS / 32);
    if (smem_tab) {
        // CDL_SELL_ASYNC=1: the cp.async pipeline, where its staging slots fit behind the weight table (measured equal to
        // plain loads at C4 and 1-4 % slower on small shards: the loop is bound by its arithmetic, not by load latency)
        constexpr bool CAN_ASYNC = KP <= 16 && sizeof(VIdx) == 2 && DEPTH == 1;
        static const bool async_on = [] { const char* e = getenv("CDL_SELL_ASYNC"); return e && e[0] == '1'; }();
        const size_t smem_async = smem + (size_t)(threads / 32) * SELL_STAGE_BYTES;
        if (CAN_ASYNC && async_on && P.parts == 1 && smem_async <= 227 * 1024) {
            if constexpr (CAN_ASYNC) {
                auto kern = k_cells_sell<KP, VIdx, true, THREADS, DEPTH, false, SKIP0, true>;
                CDL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_async));
                launch_dep(kern, dim3(grid), dim3(threads), smem_async, st, P);
                return;
            }
        }
        auto kern = P.parts > 1 ? k_cells_sell<KP, VIdx, true, THREADS, DEPTH, true, false>
                                : k_cells_sell<KP, VIdx, true, THREADS, DEPTH, false, SKIP0>;
        CDL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_dep(kern, dim3(grid), dim3(threads), smem, st, P);
    } else if (P.parts > 1) {
        launch_dep(k_cells_sell<KP, VIdx, false, THREADS, DEPTH, true, false>, dim3(grid), dim3(threads), 0, st, P);
    } else {
        launch_dep(k_cells_sell<KP, VIdx, false, THREADS, DEPTH, false, false>, dim3(grid), dim3(threads), 0, st, P);
    }
}

template <int KP>
void launch_sell_v(const Donor& d, const SellParams& P, bool smem_tab, size_t smem, int grid, int shape, bool skip0,
                   cudaStream_t st) {
    if (d.vidx32) { launch_sell_t<KP, uint32_t, 512, 2, false>(P, smem_tab, smem, grid, st); return; }
    // the base-clone shortcut is compiled for the shape large donors run in (K <= 16: the mask rides in the weights)
    // (K = 16 with the base-clone shortcut: 28 warps at 72 registers measured 1.5 % faster than 24 at 80)
    if (shape == 1 && skip0 && KP == 16) launch_sell_t<16, uint16_t, 896, 1, true>(P, smem_tab, smem, grid, st);
    else if (shape == 1 && skip0 && KP <= 16) launch_sell_t<KP, uint16_t, 768, 1, true>(P, smem_tab, smem, grid, st);
    else if (shape == 1) launch_sell_t<KP, uint16_t, 768, 1, false>(P, smem_tab, smem, grid, st);
    else launch_sell_t<KP, uint16_t, 512, 2, false>(P, smem_tab, smem, grid, st);
}

}  // namespace

void donor_free_sell(Donor& d) {
    cudaFree(d.s_off); cudaFree(d.s_perm); cudaFree(d.s_vidx); cudaFree(d.s_a); cudaFree(d.s_b); cudaFree(d.s_assign);
    d.s_off = nullptr; d.s_perm = nullptr; d.s_vidx = nullptr; d.s_a = nullptr; d.s_b = nullptr; d.s_assign = nullptr;
    d.s_assign_grid = 0; d.s_assign_warps = 0;
    d.s_slices = 0; d.s_slots = 0;
}

// Build the slice-major copy from the row layout (padding groups there are already (variant 0, 0, 0)).
void donor_build_sell(Donor& d, cudaStream_t st) {
    if (d.s_off) return;
    const int64_t M = d.M;
    const int64_t n_slices = (M + 31) / 32, slots = n_slices * 32;
    uint64_t *keys = nullptr, *keys2 = nullptr;
    uint32_t *vals = nullptr, *len = nullptr;
    void* tmp = nullptr;
    auto cleanup = [&]() { cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(tmp); };
    try {
        CDL_CUDA(cudaMalloc(&keys, 8 * (size_t)slots));
        CDL_CUDA(cudaMalloc(&keys2, 8 * (size_t)slots));
        CDL_CUDA(cudaMalloc(&vals, 4 * (size_t)slots));
        CDL_CUDA(cudaMalloc(&d.s_perm, 4 * (size_t)slots));
        const int threads = 256;
        k_sell_keys<<<(unsigned)((slots + threads - 1) / threads), threads, 0, st>>>(d.c_rowgrp, M, slots, keys, vals);
        CDL_CUDA(cudaGetLastError());
        size_t tmp_bytes = 0;
        CDL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, d.s_perm, (int)slots, 0, 64, st));
        CDL_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
        CDL_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, d.s_perm, (int)slots, 0, 64, st));
        CDL_CUDA(cudaMalloc(&d.s_off, 4 * (size_t)(n_slices + 1)));
        len = d.s_off;
        k_sell_slice_len<<<(unsigned)(((n_slices + 1) * 32 + threads - 1) / threads), threads, 0, st>>>(
            d.c_rowgrp, d.s_perm, n_slices, len);
        CDL_CUDA(cudaGetLastError());
        cudaFree(tmp); tmp = nullptr;
        tmp_bytes = 0;
        CDL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len, len, (int)(n_slices + 1), st));
        CDL_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
        CDL_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len, len, (int)(n_slices + 1), st));
        uint32_t total = 0;
        CDL_CUDA(cudaMemcpyAsync(&total, d.s_off + n_slices, 4, cudaMemcpyDeviceToHost, st));
        CDL_CUDA(cudaStreamSynchronize(st));
        const size_t words = (size_t)std::max<int64_t>((int64_t)total * 32, 1);
        const int vw = d.vidx32 ? 2 : 1;
        // variant ids below 2^14 (every donor whose weight table fits shared memory): the odd entries of a group are
        // stored pre-shifted by two, which makes their table address a single instruction (sell_group_fast)
        d.s_hi_shift = (!d.vidx32 && d.N <= 13824) ? 2 : 0;
        if ((uint64_t)words >= (1ull << 32)) throw Error(5, "slice layout: more than 2^32 group slots");
        CDL_CUDA(cudaMalloc(&d.s_vidx, 16 * words * vw));
        CDL_CUDA(cudaMalloc(&d.s_a, 16 * words));
        CDL_CUDA(cudaMalloc(&d.s_b, 16 * words));
        const unsigned blocks = (unsigned)((n_slices * 32 + threads - 1) / threads);
        if (d.vidx32)
            k_sell_fill<uint32_t><<<blocks, threads, 0, st>>>(d.c_rowgrp, d.s_perm, d.s_off, n_slices, d.N,
                (const uint32_t*)d.c_vidx, d.c_a, d.c_b, (uint32_t*)d.s_vidx, d.s_a, d.s_b, 0);
        else
            k_sell_fill<uint16_t><<<blocks, threads, 0, st>>>(d.c_rowgrp, d.s_perm, d.s_off, n_slices, d.N,
                (const uint16_t*)d.c_vidx, d.c_a, d.c_b, (uint16_t*)d.s_vidx, d.s_a, d.s_b, d.s_hi_shift);
        CDL_CUDA(cudaGetLastError());
        CDL_CUDA(cudaStreamSynchronize(st));
        d.s_slices = n_slices;
        d.s_slots = total;
    } catch (...) {
        cleanup();
        donor_free_sell(d);
        throw;
    }
    cleanup();
}

void launch_cells_sell(const Donor& d, const SweepArgs& a, ChainDev& c, int sm_count, cudaStream_t st) {
    SellParams P{};
    CellsParams& C = P.c;
    C.rowgrp = nullptr; C.vidx = d.s_vidx; C.ca = nullptr; C.cb = nullptr;
    C.M = d.M; C.N = d.N; C.cell_offset = d.cell_offset; C.K = a.K;
    C.n_tab = a.wise == 1 ? (int)a.N : 1;
    C.l1 = c.l1; C.l1c = c.l1c; C.scal = c.scal; C.mask = c.mask;
    for (int k = 0; k < KMAX; ++k) C.logPsi[k] = a.logPsi[k];
    const int64_t row = (int64_t)a.it - 1;
    C.prob_row = c.prob_all ? c.prob_all + row * d.M * a.K : nullptr;
    C.prob_acc = c.prob_acc;
    C.prob_sq = c.prob_sq;
    C.z = c.z;
    C.assign_row = c.assign_all ? c.assign_all + row * d.M : nullptr;
    C.rp_u = a.rp_u_assign;
    C.key = a.key; C.chain = a.chain; C.it = a.rng_it;
    C.chg_list = a.incremental ? c.chg_list : nullptr; C.chg_old = c.chg_old; C.chg_count = c.chg_ctl;
    C.tl = c.tl;
    P.slice_off = d.s_off; P.perm = d.s_perm;
    P.v8 = nullptr; P.a8 = (const uint4*)d.s_a; P.b8 = (const uint4*)d.s_b;
    P.n_slices = d.s_slices;
    P.hi_shift = d.s_hi_shift;
    P.work = c.work;
    // too few slices to give every SM a CTA's worth of warps: split each slice's groups over 2 or 4 warps
    {
        const int64_t per_sm = (d.s_slices + sm_count - 1) / sm_count;
        P.parts = per_sm >= 20 ? 1 : (per_sm >= 10 ? 2 : 4);
        if (const char* e = getenv("CDL_SELL_PARTS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) P.parts = v; }
        if (P.parts > 1 && (!c.sell_scratch || !c.sell_tickets)) P.parts = 1;
        P.scratch = c.sell_scratch;
        P.tickets = c.sell_tickets;
    }
    const int KP = a.K <= 4 ? 4 : (a.K <= 8 ? 8 : (a.K <= 16 ? 16 : 32));
    const bool packed = KP <= 16;
    const size_t smem = (size_t)d.N * (sizeof(double2) + (packed ? 0 : sizeof(uint32_t)));
    const bool smem_tab = smem <= 216 * 1024;
    C.tab_in_smem = smem_tab;
    // c.work is zero on entry: k_table's last block re-arms it (and the chain starts with it zeroed)
    if (!smem_tab) {
        if (!c.tab_g) throw Error(4, "global weight table not allocated");
        k_build_tab_sell<<<(unsigned)((d.N + 255) / 256), 256, 0, st>>>(d.N, C.n_tab, packed ? 1 : 0, c.l1, c.l1c,
                                                                         c.scal, c.mask, c.tab_g);
        C.l1 = reinterpret_cast<const double*>(c.tab_g);
    }
    // 24 warps per SM with one group slot of look-ahead beat 16 warps with two (the kernel is bound by the
    // fp64 pipe and needs the extra warps to keep it fed); CDL_SELL_SHAPE overrides for experiments
    // On a shard so small that even one round of slices cannot fill 16 warps per SM the extra warps buy nothing:
    // take the 16-warp shape, whose two slots of look-ahead hide the HBM latency that few warps cannot.  (A 1/8
    // shard of C4 has 27 slices per SM: one round of 27 warps, 0.122 ms, beats two rounds of 14, 0.134 ms.)
    int shape = 1;
    if (sell_warps(d.s_slices * P.parts, sm_count, (KP == 16 && a.skip0) ? 28 : 24) <= 16) shape = 0;
    if (const char* e = getenv("CDL_SELL_SHAPE")) shape = atoi(e);
    int64_t grid = sm_count;
    const int threads = d.vidx32 ? 512 : (shape == 1 ? ((KP == 16 && a.skip0) ? 896 : 768) : 512);
    // every SM gets a CTA as long as there is a work item for it: sell_warps() then sizes the CTAs so that the items fill
    // them evenly (a 1/8 shard of C4: 148 CTAs of 27 warps, not 140 full ones beside 8 idle SMs)
    const int64_t n_items = d.s_slices * P.parts;
    if (grid > n_items) grid = std::max<int64_t>(n_items, 1);
    // one-round shard: balanced static assignment of the slices to the (SM, sub-partition) bins, cached with the layout
    P.assign = nullptr;
    if (P.parts == 1) {
        const int warps = sell_warps(d.s_slices, (int)grid, threads / 32);
        if (d.s_slices <= grid * warps && d.s_slices > 0) {
            Donor& dm = const_cast<Donor&>(d);
            if (!dm.s_assign || dm.s_assign_grid != (int)grid || dm.s_assign_warps != warps) {
                std::vector<uint32_t> off((size_t)d.s_slices + 1);
                CDL_CUDA(cudaMemcpyAsync(off.data(), d.s_off, 4 * off.size(), cudaMemcpyDeviceToHost, st));
                CDL_CUDA(cudaStreamSynchronize(st));
                std::vector<int32_t> order((size_t)d.s_slices);
                for (int64_t q = 0; q < d.s_slices; ++q) order[(size_t)q] = (int32_t)q;
                std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
                    return off[(size_t)x + 1] - off[(size_t)x] > off[(size_t)y + 1] - off[(size_t)y];
                });
                // longest slice first, each into the least loaded bin that still has a free warp
                const int bins = (int)grid * 4;
                std::vector<int> used((size_t)bins, 0);
                typedef std::pair<uint64_t, int> LB;                 // (group slots so far, bin)
                std::priority_queue<LB, std::vector<LB>, std::greater<LB>> heap;
                auto cap = [&](int bin) { const int q = bin & 3; return (warps - q + 3) / 4; };
                for (int b = 0; b < bins; ++b) if (cap(b) > 0) heap.push(LB(0, b));
                std::vector<int32_t> assign((size_t)grid * warps, -1);
                for (int32_t sl : order) {
                    LB top = heap.top(); heap.pop();
                    const int bin = top.second, k = used[(size_t)bin]++;
                    assign[(size_t)(bin >> 2) * warps + (size_t)((bin & 3) + 4 * k)] = sl;
                    top.first += off[(size_t)sl + 1] - off[(size_t)sl];
                    if (used[(size_t)bin] < cap(bin)) heap.push(top);
                }
                cudaFree(dm.s_assign);
                dm.s_assign = nullptr;
                CDL_CUDA(cudaMalloc(&dm.s_assign, sizeof(int32_t) * assign.size()));
                CDL_CUDA(cudaMemcpyAsync(dm.s_assign, assign.data(), sizeof(int32_t) * assign.size(), cudaMemcpyHostToDevice, st));
                CDL_CUDA(cudaStreamSynchronize(st));
                dm.s_assign_grid = (int)grid; dm.s_assign_warps = warps;
            }
            P.assign = d.s_assign;
        }
    }
    switch (KP) {
        case 4: launch_sell_v<4>(d, P, smem_tab, smem, (int)grid, shape, a.skip0 != 0, st); break;
        case 8: launch_sell_v<8>(d, P, smem_tab, smem, (int)grid, shape, a.skip0 != 0, st); break;
        case 16: launch_sell_v<16>(d, P, smem_tab, smem, (int)grid, shape, a.skip0 != 0, st); break;
        default: launch_sell_v<32>(d, P, smem_tab, smem, (int)grid, shape, false, st); break;
    }
    CDL_CUDA(cudaGetLastError());
}

}  // namespace cdl
