// Device-side construction of the two HBM layouts of a donor from raw CSC counts.
// Replaces the reference's densification + NA normalisation (R/clone_id.R:121-122,380-391):
// an entry exists iff D > 0; A is the alternative-allele count, B = D - A the reference-allele count,
// W_log = sum lchoose(D, A).
#include "internal.h"
#include <cub/cub.cuh>

namespace cdl {

namespace {

__global__ void k_row_groups(const int64_t* __restrict__ p, int64_t M, uint32_t* __restrict__ grp) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < M) grp[j] = (uint32_t)((p[j + 1] - p[j] + GROUP - 1) / GROUP);
    if (j == M) grp[j] = 0;
}

// flags: bit0 = d == 0 present, bit1 = a > d present, bit2 = variant index out of range
template <typename VIdx>
__global__ void k_fill_cells(const int64_t* __restrict__ p, const int32_t* __restrict__ vi,
                             const uint16_t* __restrict__ a, const uint16_t* __restrict__ dd, int64_t M,
                             int64_t N, const uint32_t* __restrict__ rowgrp, VIdx* __restrict__ o_v,
                             uint16_t* __restrict__ o_a, uint16_t* __restrict__ o_b,
                             double* __restrict__ wlog_cell, unsigned long long* __restrict__ vtot,
                             unsigned long long* __restrict__ celltot_max, unsigned int* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < M; j += nwarps) {
        const int64_t e0 = p[j], e1 = p[j + 1];
        const int64_t o0 = (int64_t)rowgrp[j] * GROUP;
        double w = 0.0;
        unsigned long long tot = 0;
        unsigned int f = 0;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int32_t i = vi[e];
            const uint32_t av = a[e], dv = dd[e];
            if (dv == 0) f |= 1u;
            if (av > dv) f |= 2u;
            if (i < 0 || i >= N) { f |= 4u; continue; }
            const uint32_t bv = dv >= av ? dv - av : 0;
            o_v[o0 + (e - e0)] = (VIdx)i;
            o_a[o0 + (e - e0)] = (uint16_t)av;
            o_b[o0 + (e - e0)] = (uint16_t)bv;
            w += lgamma((double)dv + 1.0) - lgamma((double)av + 1.0) - lgamma((double)bv + 1.0);
            tot += dv;
            atomicAdd(&vtot[i], (unsigned long long)dv);
        }
        w = warp_sum_d(w);
        tot = warp_sum_u64(tot);
        f = __reduce_or_sync(0xffffffffu, f);
        if (lane == 0) {
            wlog_cell[j] = w;
            atomicMax(celltot_max, tot);
            if (f) atomicOr(flags, f);
        }
    }
}

__global__ void k_seg_count(const int64_t* __restrict__ p, const int32_t* __restrict__ vi,
                            const uint16_t* __restrict__ dd, int64_t M, int64_t N, uint32_t* __restrict__ cnt,
                            uint32_t* __restrict__ maxdepth) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < M; j += nwarps) {
        const int64_t e0 = p[j], e1 = p[j + 1];
        const int64_t cb = j / CELL_BLOCK;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int32_t i = vi[e];
            if (i >= 0 && i < N) {
                atomicAdd(&cnt[cb * N + i], 1u);
                atomicMax(&maxdepth[cb * N + i], (uint32_t)dd[e]);
            }
        }
    }
}

__global__ void k_to_groups(uint32_t* __restrict__ cnt, int64_t n, const uint32_t* __restrict__ maxdepth,
                            uint16_t* __restrict__ flush) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s < n) {
        cnt[s] = (cnt[s] + GROUP - 1) / GROUP;
        // k_stats keeps 16 + 16 bit packed bins per lane: a lane may add this many 8-entry groups before
        // either half could overflow (0 = even one group could: that segment takes the 64-bit path)
        const uint32_t md = maxdepth[s] ? maxdepth[s] : 1u;
        const uint32_t f = 65535u / (GROUP * md);
        flush[s] = (uint16_t)(f > 65535u ? 65535u : f);
    }
    if (s == n) cnt[s] = 0;
}

__global__ void k_fill_segs(const int64_t* __restrict__ p, const int32_t* __restrict__ vi,
                            const uint16_t* __restrict__ a, const uint16_t* __restrict__ dd, int64_t M,
                            int64_t N, const uint32_t* __restrict__ seggrp, uint32_t* __restrict__ cursor,
                            uint16_t* __restrict__ o_c, uint16_t* __restrict__ o_a,
                            uint16_t* __restrict__ o_b) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < M; j += nwarps) {
        const int64_t e0 = p[j], e1 = p[j + 1];
        const int64_t cb = j / CELL_BLOCK;
        const uint16_t cl = (uint16_t)(j - cb * CELL_BLOCK);
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int32_t i = vi[e];
            if (i < 0 || i >= N) continue;
            const int64_t s = cb * N + i;
            const uint32_t slot = atomicAdd(&cursor[s], 1u);
            const int64_t o = (int64_t)seggrp[s] * GROUP + slot;
            const uint32_t av = a[e], dv = dd[e];
            o_c[o] = cl;
            o_a[o] = (uint16_t)av;
            o_b[o] = (uint16_t)(dv >= av ? dv - av : 0);
        }
    }
}

template <typename T>
T* dalloc(int64_t n) {
    T* ptr = nullptr;
    if (n <= 0) n = 1;
    cudaError_t e = cudaMalloc(&ptr, sizeof(T) * (size_t)n);
    if (e != cudaSuccess) throw Error(3, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    return ptr;
}

void exclusive_scan_u32(uint32_t* data, int64_t n, cudaStream_t st) {
    size_t tmp_bytes = 0;
    CDL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, data, data, (int)n, st));
    void* tmp = dalloc<uint8_t>((int64_t)tmp_bytes);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, data, data, (int)n, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    CDL_CUDA(e);
}

double sum_doubles(const double* data, int64_t n, cudaStream_t st) {
    if (n == 0) return 0.0;
    double* out = dalloc<double>(1);
    size_t tmp_bytes = 0;
    CDL_CUDA(cub::DeviceReduce::Sum(nullptr, tmp_bytes, data, out, (int)n, st));
    void* tmp = dalloc<uint8_t>((int64_t)tmp_bytes);
    cudaError_t e = cub::DeviceReduce::Sum(tmp, tmp_bytes, data, out, (int)n, st);
    double h = 0.0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, out, sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    cudaFree(out);
    CDL_CUDA(e);
    return h;
}

}  // namespace

void donor_free(Donor& d) {
    donor_free_sell(d);
    cudaFree(d.c_elem0);
    cudaFree(d.c_rowgrp); cudaFree(d.c_vidx); cudaFree(d.c_a); cudaFree(d.c_b);
    cudaFree(d.v_seggrp); cudaFree(d.v_flush); cudaFree(d.v_plan); cudaFree(d.v_cell); cudaFree(d.v_a); cudaFree(d.v_b);
    cudaFree(d.labels);
    d = Donor();
}

void donor_build(Donor& d, int64_t N, int64_t M, int64_t nnz, const int64_t* p, const int32_t* vi,
                 const uint16_t* a, const uint16_t* dd, cudaStream_t st) {
    int32_t* keep_labels = d.labels;
    d.labels = nullptr;
    int64_t keep_Mt = d.M_total, keep_off = d.cell_offset;
    donor_free(d);
    d.labels = keep_labels;
    d.N = N; d.M = M; d.nnz = nnz;
    d.M_total = keep_Mt > 0 ? keep_Mt : M;
    d.cell_offset = keep_off;
    d.vidx32 = N > 65536;
    if (N <= 0 || M <= 0) throw Error(1, "A and D must have at least one variant and one cell");
    if (nnz / GROUP + M >= (int64_t)0xFFFFFFF0ll) throw Error(5, "too many entries for 32-bit group offsets");

    const int threads = 256;
    // ---- layout (A) ----
    d.c_rowgrp = dalloc<uint32_t>(M + 1);
    k_row_groups<<<(unsigned)((M + 1 + threads - 1) / threads), threads, 0, st>>>(p, M, d.c_rowgrp);
    exclusive_scan_u32(d.c_rowgrp, M + 1, st);
    uint32_t total_groups = 0;
    CDL_CUDA(cudaMemcpyAsync(&total_groups, d.c_rowgrp + M, 4, cudaMemcpyDeviceToHost, st));
    CDL_CUDA(cudaStreamSynchronize(st));
    d.c_groups = total_groups;
    const int64_t ne = (int64_t)total_groups * GROUP;
    const size_t vbytes = (size_t)(d.vidx32 ? 4 : 2) * (size_t)(ne > 0 ? ne : 1);
    CDL_CUDA(cudaMalloc(&d.c_vidx, vbytes));
    d.c_a = dalloc<uint16_t>(ne);
    d.c_b = dalloc<uint16_t>(ne);
    CDL_CUDA(cudaMemsetAsync(d.c_vidx, 0, vbytes, st));
    CDL_CUDA(cudaMemsetAsync(d.c_a, 0, sizeof(uint16_t) * (size_t)(ne > 0 ? ne : 1), st));
    CDL_CUDA(cudaMemsetAsync(d.c_b, 0, sizeof(uint16_t) * (size_t)(ne > 0 ? ne : 1), st));

    double* wlog_cell = dalloc<double>(M);
    unsigned long long* vtot = dalloc<unsigned long long>(N + 2);
    unsigned int* flags = dalloc<unsigned int>(1);
    CDL_CUDA(cudaMemsetAsync(vtot, 0, sizeof(unsigned long long) * (size_t)(N + 2), st));
    CDL_CUDA(cudaMemsetAsync(flags, 0, 4, st));
    const unsigned blocks = (unsigned)std::min<int64_t>((M * 32 + threads - 1) / threads, 148 * 16);
    if (d.vidx32)
        k_fill_cells<uint32_t><<<blocks, threads, 0, st>>>(p, vi, a, dd, M, N, d.c_rowgrp, (uint32_t*)d.c_vidx,
                                                           d.c_a, d.c_b, wlog_cell, vtot, vtot + N, flags);
    else
        k_fill_cells<uint16_t><<<blocks, threads, 0, st>>>(p, vi, a, dd, M, N, d.c_rowgrp, (uint16_t*)d.c_vidx,
                                                           d.c_a, d.c_b, wlog_cell, vtot, vtot + N, flags);
    CDL_CUDA(cudaGetLastError());
    d.W_log = sum_doubles(wlog_cell, M, st);
    unsigned int hflags = 0;
    CDL_CUDA(cudaMemcpyAsync(&hflags, flags, 4, cudaMemcpyDeviceToHost, st));
    {
        // max per-variant depth via a tiny host pass (N values)
        std::string tmp;
        unsigned long long* hv = (unsigned long long*)malloc(sizeof(unsigned long long) * (size_t)(N + 1));
        cudaError_t e = cudaMemcpyAsync(hv, vtot, sizeof(unsigned long long) * (size_t)(N + 1),
                                        cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        unsigned long long mx = 0;
        if (e == cudaSuccess)
            for (int64_t i = 0; i < N; ++i) mx = hv[i] > mx ? hv[i] : mx;
        d.max_variant_depth = mx;
        d.max_cell_depth = hv[N];
        free(hv);
        CDL_CUDA(e);
    }
    cudaFree(wlog_cell); cudaFree(vtot); cudaFree(flags);
    if (hflags & 4u) throw Error(1, "variant index out of range in sparse input");
    if (hflags & 1u) throw Error(1, "explicit zero depth in compact sparse input");
    if (hflags & 2u) throw Error(1, "A exceeds D for some entry (alternative reads > total reads)");

    // ---- layout (B) ----
    d.n_cb = (int)((M + CELL_BLOCK - 1) / CELL_BLOCK);
    const int64_t nseg = (int64_t)d.n_cb * N;
    if (nseg + 1 >= (int64_t)0x7FFFFFFF) throw Error(5, "too many variant segments");
    d.v_seggrp = dalloc<uint32_t>(nseg + 1);
    uint32_t* cursor = dalloc<uint32_t>(nseg + 1);
    CDL_CUDA(cudaMemsetAsync(d.v_seggrp, 0, 4 * (size_t)(nseg + 1), st));
    CDL_CUDA(cudaMemsetAsync(cursor, 0, 4 * (size_t)(nseg + 1), st));
    uint32_t* segdepth = dalloc<uint32_t>(nseg + 1);
    CDL_CUDA(cudaMemsetAsync(segdepth, 0, 4 * (size_t)(nseg + 1), st));
    d.v_flush = dalloc<uint16_t>(nseg + 1);
    k_seg_count<<<blocks, threads, 0, st>>>(p, vi, dd, M, N, d.v_seggrp, segdepth);
    k_to_groups<<<(unsigned)((nseg + 1 + threads - 1) / threads), threads, 0, st>>>(d.v_seggrp, nseg, segdepth,
                                                                                      d.v_flush);
    exclusive_scan_u32(d.v_seggrp, nseg + 1, st);
    uint32_t vg = 0;
    CDL_CUDA(cudaMemcpyAsync(&vg, d.v_seggrp + nseg, 4, cudaMemcpyDeviceToHost, st));
    CDL_CUDA(cudaStreamSynchronize(st));
    d.v_groups = vg;
    const int64_t nv = (int64_t)vg * GROUP;
    d.v_cell = dalloc<uint16_t>(nv);
    d.v_a = dalloc<uint16_t>(nv);
    d.v_b = dalloc<uint16_t>(nv);
    CDL_CUDA(cudaMemsetAsync(d.v_cell, 0, 2 * (size_t)(nv > 0 ? nv : 1), st));
    CDL_CUDA(cudaMemsetAsync(d.v_a, 0, 2 * (size_t)(nv > 0 ? nv : 1), st));
    CDL_CUDA(cudaMemsetAsync(d.v_b, 0, 2 * (size_t)(nv > 0 ? nv : 1), st));
    k_fill_segs<<<blocks, threads, 0, st>>>(p, vi, a, dd, M, N, d.v_seggrp, cursor, d.v_cell, d.v_a, d.v_b);
    CDL_CUDA(cudaGetLastError());
    CDL_CUDA(cudaStreamSynchronize(st));
    cudaFree(cursor);
    cudaFree(segdepth);
}

namespace {
template <typename VIdx>
__global__ void k_export(const uint32_t* __restrict__ rowgrp, const VIdx* __restrict__ v,
                         const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, int64_t M,
                         const int64_t* __restrict__ p, int32_t* __restrict__ oi, uint16_t* __restrict__ oa,
                         uint16_t* __restrict__ od) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < M; j += nwarps) {
        const int64_t o0 = p[j], len = p[j + 1] - p[j];
        const int64_t s0 = (int64_t)rowgrp[j] * GROUP;
        for (int64_t e = lane; e < len; e += 32) {
            oi[o0 + e] = (int32_t)v[s0 + e];
            oa[o0 + e] = a[s0 + e];
            od[o0 + e] = (uint16_t)(a[s0 + e] + b[s0 + e]);
        }
    }
}

// true (unpadded) row lengths: padding entries have a = b = 0 and real entries have a + b > 0
__global__ void k_row_len(const uint32_t* __restrict__ rowgrp, const uint16_t* __restrict__ a,
                          const uint16_t* __restrict__ b, int64_t M, int64_t* __restrict__ len) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int64_t s0 = (int64_t)rowgrp[j] * GROUP, s1 = (int64_t)rowgrp[j + 1] * GROUP;
    int64_t n = s1 - s0;
    while (n > 0 && a[s0 + n - 1] == 0 && b[s0 + n - 1] == 0) --n;
    len[j] = n;
}
}  // namespace

// element index of each cell's first covered entry (prefix sums of the true row lengths)
void donor_build_elem0(Donor& d, cudaStream_t st) {
    if (d.c_elem0) return;
    int64_t* len = dalloc<int64_t>(d.M + 1);
    CDL_CUDA(cudaMemsetAsync(len, 0, sizeof(int64_t) * (size_t)(d.M + 1), st));
    const int threads = 256;
    k_row_len<<<(unsigned)((d.M + threads - 1) / threads), threads, 0, st>>>(d.c_rowgrp, d.c_a, d.c_b, d.M, len);
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len, len, (int)(d.M + 1), st);
    void* tmp = dalloc<uint8_t>((int64_t)tmp_bytes);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len, len, (int)(d.M + 1), st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) { cudaFree(len); CDL_CUDA(e); }
    d.c_elem0 = len;
}

void donor_export(const Donor& d, int64_t* p, int32_t* oi, uint16_t* oa, uint16_t* od, cudaStream_t st) {
    int64_t* len = dalloc<int64_t>(d.M + 1);
    const int threads = 256;
    k_row_len<<<(unsigned)((d.M + threads - 1) / threads), threads, 0, st>>>(d.c_rowgrp, d.c_a, d.c_b,
                                                                                      d.M, len);
    std::string err;
    int64_t* hl = (int64_t*)malloc(sizeof(int64_t) * (size_t)(d.M + 1));
    CDL_CUDA(cudaMemcpyAsync(hl, len, sizeof(int64_t) * (size_t)d.M, cudaMemcpyDeviceToHost, st));
    CDL_CUDA(cudaStreamSynchronize(st));
    p[0] = 0;
    for (int64_t j = 0; j < d.M; ++j) p[j + 1] = p[j] + hl[j];
    free(hl);
    int64_t* p_dev = dalloc<int64_t>(d.M + 1);
    CDL_CUDA(cudaMemcpyAsync(p_dev, p, sizeof(int64_t) * (size_t)(d.M + 1), cudaMemcpyHostToDevice, st));
    const int64_t nnz = p[d.M];
    int32_t* di = dalloc<int32_t>(nnz);
    uint16_t* da = dalloc<uint16_t>(nnz);
    uint16_t* ddd = dalloc<uint16_t>(nnz);
    const unsigned blocks = (unsigned)std::min<int64_t>((d.M * 32 + threads - 1) / threads, 148 * 16);
    if (d.vidx32)
        k_export<uint32_t><<<blocks, threads, 0, st>>>(d.c_rowgrp, (const uint32_t*)d.c_vidx, d.c_a, d.c_b, d.M,
                                                       p_dev, di, da, ddd);
    else
        k_export<uint16_t><<<blocks, threads, 0, st>>>(d.c_rowgrp, (const uint16_t*)d.c_vidx, d.c_a, d.c_b, d.M,
                                                       p_dev, di, da, ddd);
    CDL_CUDA(cudaGetLastError());
    CDL_CUDA(cudaMemcpyAsync(oi, di, 4 * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    CDL_CUDA(cudaMemcpyAsync(oa, da, 2 * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    CDL_CUDA(cudaMemcpyAsync(od, ddd, 2 * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    CDL_CUDA(cudaStreamSynchronize(st));
    cudaFree(len); cudaFree(p_dev); cudaFree(di); cudaFree(da); cudaFree(ddd);
}

}  // namespace cdl
