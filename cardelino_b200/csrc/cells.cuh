// Device pieces shared by the two cell kernels (row-split k_cells in gibbs.cu, slice-major k_cells_sell
// in cells_sell.cu): parameters, the masked fp64 accumulate of one 8-entry group (R/clone_id.R:470-478) and
// R's revsort for exactly tied probabilities (base R src/main/sort.c, used by sample() at :493).
#pragma once
#include "internal.h"

namespace cdl {
namespace {

template <int KP> struct Log2 {};
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };
template <> struct Log2<16> { static constexpr int v = 4; };
template <> struct Log2<32> { static constexpr int v = 5; };

struct CellsParams {
    const uint32_t* rowgrp;
    const void* vidx;
    const uint16_t* ca;
    const uint16_t* cb;
    int64_t M, N, cell_offset;
    int K;
    int n_tab;                 // N (variant) or 1 (global)
    const double* l1;
    const double* l1c;
    const double* scal;
    const uint32_t* mask;
    double logPsi[KMAX];
    double* prob_row;          // [M*K] trace row to fill (device layout [cell][clone]) or nullptr
    double* prob_acc;          // [M*K] accumulator or nullptr
    double* prob_sq;           // [M*K] accumulator of squares (streaming Geweke windows) or nullptr
    uint8_t* z;
    uint8_t* assign_row;       // [M] 1-based labels or nullptr
    const double* rp_u;        // replay uniforms for this iteration or nullptr
    uint32_t* chg_list;        // incremental statistics: cells whose label changes are appended here (or nullptr)
    uint8_t* chg_old;
    unsigned int* chg_count;
    PhiloxKey key;
    uint32_t chain, it;
    int tab_in_smem;
    int zero;                  // always 0; a kernel parameter so it lives in a register (see accumulate_entry)
    unsigned long long* tl;    // device timeline (common.cuh::tl_stamp) or nullptr
};

// R's revsort (src/main/sort.c) on a tiny array: used only when positive probabilities tie exactly,
// where sample()'s result depends on the heapsort's tie order.
__device__ void revsort_dev(double* a, int* ib, int n) {
    if (n <= 1) return;
    int l = (n >> 1) + 1, ir = n;
    double ra;
    int ii;
    for (;;) {
        if (l > 1) {
            l = l - 1;
            ra = a[l - 1];
            ii = ib[l - 1];
        } else {
            ra = a[ir - 1];
            ii = ib[ir - 1];
            a[ir - 1] = a[0];
            ib[ir - 1] = ib[0];
            if (--ir == 1) {
                a[0] = ra;
                ib[0] = ii;
                return;
            }
        }
        int i = l, j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j - 1] > a[j]) ++j;
            if (ra > a[j - 1]) {
                a[i - 1] = a[j - 1];
                ib[i - 1] = ib[j - 1];
                i = j;
                j += i;
            } else {
                j = ir + 1;
            }
        }
        a[i - 1] = ra;
        ib[i - 1] = ii;
    }
}

// acc[k] += bit_k(m) ? del : 0 for all clones, as acc[k] = fma(del, w_k, acc[k]) with w_k in {0.0, 1.0}
// built from one SEL on the high word (ptxas turns a predicated fp64 add into two FSELs plus an
// unconditional DADD, and produces the predicates seven at a time with R2P): one SEL + one DFMA per
// clone.  `zero` must be a register holding 0 (the low word of every weight).
template <int KP>
__device__ __forceinline__ void accumulate_entry(double (&acc)[KP], double del, uint32_t m, int zero) {
#ifndef CDL_ACC_VARIANT
#define CDL_ACC_VARIANT 3
#endif
#if CDL_ACC_VARIANT == 1
    (void)zero;
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] += ((m >> k) & 1u) ? del : 0.0;
#elif CDL_ACC_VARIANT == 3
    // hi-word masking: (lo, hi & -bit) is either del or a denormal < 2^-1022 that any normal accumulator
    // absorbs; one SEL (alu pipe) + one DADD (fp64 pipe) per clone instead of two FSELs + DADD.
    (void)zero;
    const int hi = __double2hiint(del), lo = __double2loint(del);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, h;\n\t.reg .f64 w;\n\t"
            "and.b32 t, %3, %4;\n\tsetp.ne.u32 p, t, 0;\n\tselp.b32 h, %2, 0, p;\n\t"
            "mov.b64 w, {%1, h};\n\tadd.f64 %0, %0, w;\n\t}"
            : "+d"(acc[k])
            : "r"(lo), "r"(hi), "r"(m), "r"(1u << k));
    }
#else
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        asm("{\n\t.reg .pred p;\n\t.reg .b32 t, h;\n\t.reg .f64 w;\n\t"
            "and.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\tselp.b32 h, 0x3ff00000, 0, p;\n\t"
            "mov.b64 w, {%4, h};\n\tfma.rn.f64 %0, %1, w, %0;\n\t}"
            : "+d"(acc[k])
            : "d"(del), "r"(m), "r"(1u << k), "r"(zero));
    }
#endif
}

// weight table access: shared memory through 32-bit shared addresses, or global memory through L1
template <bool SMEM>
__device__ __forceinline__ void load_tab(uint32_t tab_sa, uint32_t msk_sa, const double2* __restrict__ tab_g,
                                         const uint32_t* __restrict__ msk_g, uint32_t i, double& du, double& dv,
                                         uint32_t& m) {
    if (SMEM) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(du), "=d"(dv) : "r"(tab_sa + (i << 4)));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(msk_sa + (i << 2)));
    } else {
        const double2 t = __ldg(tab_g + i);
        du = t.x; dv = t.y;
        m = __ldg(msk_g + i);
    }
}

// one group of 8 entries as it sits in HBM
template <typename VIdx> struct Grp;
template <> struct Grp<uint16_t> {
    uint4 v, a, b;
    __device__ __forceinline__ void load(const void* vidx, const uint4* a8, const uint4* b8, int64_t g) {
        v = ld_stream(reinterpret_cast<const uint4*>(vidx) + g);
        a = ld_stream(a8 + g);
        b = ld_stream(b8 + g);
    }
    __device__ __forceinline__ uint32_t idx(int e) const {
        const uint32_t w = e < 2 ? v.x : (e < 4 ? v.y : (e < 6 ? v.z : v.w));
        return (e & 1) ? (w >> 16) : (w & 0xffffu);
    }
};
template <> struct Grp<uint32_t> {
    uint4 v, v2, a, b;
    __device__ __forceinline__ void load(const void* vidx, const uint4* a8, const uint4* b8, int64_t g) {
        v = ld_stream(reinterpret_cast<const uint4*>(vidx) + 2 * g);
        v2 = ld_stream(reinterpret_cast<const uint4*>(vidx) + 2 * g + 1);
        a = ld_stream(a8 + g);
        b = ld_stream(b8 + g);
    }
    __device__ __forceinline__ uint32_t idx(int e) const {
        return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : e == 3 ? v.w : e == 4 ? v2.x : e == 5 ? v2.y
                                                                                 : e == 6 ? v2.z : v2.w;
    }
};

template <int KP, typename VIdx, bool SMEM>
__device__ __forceinline__ void accumulate_group(double (&acc)[KP], const Grp<VIdx>& q, uint32_t tab_sa,
                                                 uint32_t msk_sa, const double2* __restrict__ tab_g,
                                                 const uint32_t* __restrict__ msk_g, int zero) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t aw = e < 2 ? q.a.x : (e < 4 ? q.a.y : (e < 6 ? q.a.z : q.a.w));
        const uint32_t bw = e < 2 ? q.b.x : (e < 4 ? q.b.y : (e < 6 ? q.b.z : q.b.w));
        const uint32_t a = (e & 1) ? (aw >> 16) : (aw & 0xffffu);
        const uint32_t b = (e & 1) ? (bw >> 16) : (bw & 0xffffu);
        double du, dv;
        uint32_t m;
        load_tab<SMEM>(tab_sa, msk_sa, tab_g, msk_g, q.idx(e), du, dv, m);
        const double del = fma(u16_to_double(a), du, u16_to_double(b) * dv);
        accumulate_entry<KP>(acc, del, m, zero);
    }
}


// Per-lane softmax over K logits held in registers (R/clone_id.R:479-481): acc <- probabilities.
template <int KP>
__device__ __forceinline__ void lane_softmax(double (&acc)[KP], int K, const double* logPsi) {
    double mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < KP; ++k)
        if (k < K) { acc[k] += logPsi[k]; mx = fmax(mx, acc[k]); }
    double sm = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (k < K) { acc[k] = exp(acc[k] - mx); sm += acc[k]; }
        else acc[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < KP; ++k)
        if (k < K) acc[k] = acc[k] / sm;
}

// sample(seq_len(K), 1, prob = pr) with the uniform u (:493): walk the DESCENDING-sorted CDF; exact ties are
// ordered by R's revsort.  Every lane of the warp must call it (warp votes end the walk early).
template <int KP>
__device__ __forceinline__ int lane_draw(const double (&pr)[KP], int K, double u, bool valid) {
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
            const bool take = (k < K) && !((used >> k) & 1u) && (pr[k] > bv);   // ties -> lowest index
            bv = take ? pr[k] : bv;
            bk = take ? k : bk;
        }
        if (pending) {
            used |= 1u << bk;
            cum += bv;
            if (u <= cum || r == K - 1) { sel = bk; selv = bv; pending = false; }
        }
    }
    int ntie = 0;
#pragma unroll
    for (int k = 0; k < KP; ++k) ntie += (k < K && pr[k] == selv) ? 1 : 0;
    if (valid && ntie > 1 && selv > 0.0) {
        double pa[KP];
        int perm[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) { pa[k] = pr[k]; perm[k] = k; }
        double tot = 0.0;
        for (int l = 0; l < K; ++l) tot += pa[l];
        for (int l = 0; l < K; ++l) pa[l] /= tot;
        revsort_dev(pa, perm, K);
        for (int l = 1; l < K; ++l) pa[l] += pa[l - 1];
        int jj = 0;
        while (jj < K - 1 && !(u <= pa[jj])) ++jj;
        sel = perm[jj];
    }
    return sel;
}

// label write-back; with incremental statistics the cells whose label changed are appended to a list (one
// atomic per warp)
__device__ __forceinline__ void store_label(const CellsParams& C, bool writer, int64_t cell, int sel) {
    if (C.chg_list) {
        const uint8_t old = writer ? C.z[cell] : (uint8_t)sel;
        const bool changed = writer && old != (uint8_t)sel;
        const unsigned m = __ballot_sync(0xffffffffu, changed);
        if (m) {
            const int lane = threadIdx.x & 31;
            unsigned base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(C.chg_count, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (changed) {
                const unsigned idx = base + (unsigned)__popc(m & ((1u << lane) - 1u));
                if ((int64_t)idx < C.M) { C.chg_list[idx] = (uint32_t)cell; C.chg_old[idx] = old; }
            }
        }
    }
    if (writer) {
        C.z[cell] = (uint8_t)sel;
        if (C.assign_row) C.assign_row[cell] = (uint8_t)(sel + 1);
    }
}

}  // namespace
}  // namespace cdl
