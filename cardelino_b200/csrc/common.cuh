// Shared device helpers: Philox4x32-10 streams, warp utilities, Beta sampling.
// The Philox keying is restated in oracle/philox.py so the checker reproduces every uniform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace cdl {

enum Stream : uint32_t {
    STREAM_ASSIGN = 0,   // categorical clone draw, idx = global cell id            (R/clone_id.R:493)
    STREAM_CONFIG = 1,   // Bernoulli Config flip, idx = i + k*N (column-major)     (R/clone_id.R:534)
    STREAM_THETA0 = 2,   // Beta draw of theta0                                     (R/clone_id.R:461,563)
    STREAM_THETA1 = 3,   // Beta draw of theta1[idx]                                (R/clone_id.R:462,567)
    STREAM_RELAX = 4,    // Beta draw of relax_rate                                 (R/clone_id.R:504)
    STREAM_EM_INIT = 5,  // runif starts of clone_id_EM                             (R/clone_id.R:278)
    STREAM_SYNTH = 16    // synthetic-donor generator (+ sub-streams)
};

struct PhiloxKey {
    uint32_t k0, k1;
};

__host__ __device__ inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                             uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two words -> double strictly inside (0,1), 53-bit resolution
__host__ __device__ inline double u53(uint32_t x0, uint32_t x1) {
    const double hi = (double)(x0 >> 5);
    const double lo = (double)(x1 >> 6);
    return (hi * 67108864.0 + lo + 0.5) * (1.0 / 9007199254740992.0);
}

__host__ __device__ inline uint32_t stream_word(uint32_t chain, uint32_t stream) {
    return (chain << 8) | stream;
}

// the one uniform of (chain, iteration, stream, idx)
__host__ __device__ inline double philox_uniform(PhiloxKey key, uint32_t chain, uint32_t it, uint32_t stream,
                                                 uint32_t idx, uint32_t sub = 0) {
    uint32_t o[4];
    philox4x32_10(idx, sub, it, stream_word(chain, stream), key.k0, key.k1, o);
    return u53(o[0], o[1]);
}

// sequential generator over `sub` for rejection samplers
struct PhiloxSeq {
    PhiloxKey key;
    uint32_t idx, it, sw, sub;
    __host__ __device__ PhiloxSeq(PhiloxKey k, uint32_t chain, uint32_t it_, uint32_t stream, uint32_t idx_,
                                  uint32_t sub0 = 0)
        : key(k), idx(idx_), it(it_), sw(stream_word(chain, stream)), sub(sub0) {}
    // two 53-bit uniforms per block
    __host__ __device__ inline void next2(double& u, double& v) {
        uint32_t o[4];
        philox4x32_10(idx, sub++, it, sw, key.k0, key.k1, o);
        u = u53(o[0], o[1]);
        v = u53(o[2], o[3]);
    }
};

// Marsaglia & Tsang (2000) Gamma(a, 1); a < 1 handled with the U^(1/a) boost.
// One attempt consumes two Philox blocks and is split in two so that the part that does not depend on the shape -- the
// normal deviate and the acceptance uniform, most of the work -- can be made before the shape is known (k_table makes
// the candidates of every draw of the sweep while it waits for the statistic kernel / the peers' tables):
//   gamma_candidate  blocks sub, sub + 1  ->  x ~ N(0, 1), u3 ~ U(0, 1)
//   gamma_accept     the squeeze and the log test for shape parameter d = a - 1/3, c = 1 / sqrt(9 d); v = (1 + c x)^3
__host__ __device__ inline void gamma_candidate(PhiloxSeq& s, double& x, double& u3) {
    double u1, u2, u4;
    s.next2(u1, u2);
    s.next2(u3, u4);
#ifdef __CUDA_ARCH__
    x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
#else
    x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
#endif
}
__host__ __device__ inline bool gamma_accept(double d, double c, double x, double u3, double& v) {
    v = 1.0 + c * x;
    if (v <= 0.0) return false;
    v = v * v * v;
    const double x2 = x * x;
    if (u3 < 1.0 - 0.0331 * x2 * x2) return true;
    return log(u3) < 0.5 * x2 + d * (1.0 - v + log(v));
}
// the attempts from the sequence's current position on (at most `max_attempts`)
__host__ __device__ inline double gamma_mt_attempts(double d, double c, PhiloxSeq& s, int max_attempts) {
    for (int attempt = 0; attempt < max_attempts; ++attempt) {
        double x, u3, v;
        gamma_candidate(s, x, u3);
        if (gamma_accept(d, c, x, u3, v)) return d * v;
    }
    return d;  // unreachable in practice (acceptance > 95 %)
}
__host__ __device__ inline double gamma_mt(double a, PhiloxSeq& s) {
    double boost = 1.0;
    if (a < 1.0) {
        double u, v;
        s.next2(u, v);
        boost = pow(u, 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    return boost * gamma_mt_attempts(d, c, s, 1000);
}

// Beta(a, b) as X / (X + Y); clamped strictly inside (0,1) so both logs stay finite
// (the reference would produce log(0) = -Inf if rbeta returned exactly 0 or 1).
__host__ __device__ inline double clamp_theta(double t) {
    const double lo = 1e-300, hi = 1.0 - 1.1102230246251565e-16;
    return t < lo ? lo : (t > hi ? hi : t);
}

// Beta(a, b) from two independent Gamma draws X ~ Gamma(a), Y ~ Gamma(b) made by different threads
__host__ __device__ inline double beta_from_gammas(double x, double y) {
    const double sum = x + y;
    return clamp_theta((sum > 0.0) ? x / sum : 0.5);
}

__host__ __device__ inline double beta_draw(double a, double b, PhiloxSeq& s) {
    const double x = gamma_mt(a, s);
    const double y = gamma_mt(b, s);
    const double sum = x + y;
    double t = (sum > 0.0) ? x / sum : 0.5;
    return clamp_theta(t);
}

// rbinom(1, 1, p) with its uniform injected (base R nmath/rbinom.c, size-1 inversion branch)
__host__ __device__ inline int rbinom1(double p, double u) {
    if (p == 0.0) return 0;
    if (p == 1.0) return 1;
    const double pm = p < 1.0 - p ? p : 1.0 - p;
    const double q = 1.0 - pm;
    int ix = (u < q) ? 0 : 1;
    if (p > 0.5) ix = 1 - ix;
    return ix;
}

#ifdef __CUDACC__
__device__ __forceinline__ double shfl_xor_d(double v, int off) {
    return __shfl_xor_sync(0xffffffffu, v, off);
}
__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// First statement of every kernel launched with launch_dep() (internal.h): let the next kernel's CTAs be scheduled
// as soon as there is room for them, then wait until the previous kernel has completed and flushed its writes.
// Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

// Device timeline (CDL_TIMELINE=1, api.cu::cdl_bench_sweeps prints it): thread 0 of a block stores %globaltimer into slot
// `slot` of its block's row; `tl` is nullptr in ordinary runs (one predictable branch per stamp).  TL_ROWS blocks of
// TL_SLOTS stamps per kernel (cell, statistic, table kernel in that order).
constexpr int TL_SLOTS = 8, TL_ROWS = 1024;
__device__ __forceinline__ void tl_stamp(unsigned long long* tl, int kernel, int slot) {
    if (tl && threadIdx.x == 0 && blockIdx.x < TL_ROWS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tl[((size_t)kernel * TL_ROWS + blockIdx.x) * TL_SLOTS + slot] = t;
    }
}
// exact u16 -> double without the (slow) I2F.F64 path: 2^52 + a has `a` in the low mantissa bits
__device__ __forceinline__ double u16_to_double(uint32_t a) {
    return __hiloint2double(0x43300000, (int)a) - 4503599627370496.0;
}
// 64-bit load of another GPU's memory over NVLink (or of local memory): system scope, never from a stale L1 line
__device__ __forceinline__ unsigned long long ld_peer_u64(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// streaming 128-bit load, bypassing L1 allocation (read-once data)
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
#endif

}  // namespace cdl
