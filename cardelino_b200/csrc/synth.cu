// Synthetic donor generated on the device with the distributions of the reference's simulator
// (R/reads_simulator.R:53-152 sim_read_count, :174-212 sample_seq_depth):
//   clone label  ~ Multinomial(1, Psi = 1/K)                       (:80)
//   genotype     H = Config[i, label]                              (:83)
//   coverage     ~ Bernoulli(cov) per entry (geometric gaps), depth resampled from the empirical
//                distribution of the packaged D_input (median 2, p90 10, p99 48, max 247)      (:197-209)
//   theta0       ~ Beta(0.002*100, 0.998*100) per entry, theta1 ~ Beta(0.45, 0.55) per variant  (:54-55,95-109)
//   A            ~ Binomial(D, theta1 if H else theta0)            (:125-129)
// Every draw is keyed by the GLOBAL cell id, so a rank's shard is identical to the same cells of the
// single-GPU donor.
#include "internal.h"
#include <vector>

namespace cdl {

namespace {

// inverse-CDF table of D_input's positive depths at quantiles (q + 0.5) / 256 (tools/rdata_to_npz.py)
__constant__ uint8_t c_depth_q[256] = {
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
    1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,2,2,2,2,2,2,2,2,
    2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,2,3,3,3,3,3,3,3,3,3,3,3,3,
    3,3,3,3,3,3,3,3,3,3,3,3,3,3,3,3,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,5,5,5,5,5,5,5,5,5,5,5,6,6,6,6,6,6,6,6,
    6,7,7,7,7,7,7,8,8,8,8,8,9,9,9,9,10,10,10,11,11,11,12,12,12,13,13,14,14,15,16,17,18,19,21,22,24,27,30,34,
    38,49,61,107};

__device__ inline uint32_t binom_inv(uint32_t n, double p, double u) {
    const bool flip = p > 0.5;
    if (flip) p = 1.0 - p;
    const double q = 1.0 - p, r = p / q;
    double f = pow(q, (double)n);
    uint32_t ix = 0;
    while (ix < n) {
        if (u < f) break;
        u -= f;
        f *= r * (double)(n - ix) / (double)(ix + 1);
        ++ix;
    }
    return flip ? n - ix : ix;
}

// FILL = false: count entries per cell.  FILL = true: write them.  Same Philox sequence in both.
template <bool FILL>
__global__ void k_synth(int64_t N, int64_t M_local, int64_t cell_offset, int K, double log1m_cov,
                        const uint32_t* __restrict__ cfg, const double* __restrict__ theta1, PhiloxKey key,
                        int64_t* __restrict__ counts, const int64_t* __restrict__ p, int32_t* __restrict__ oi,
                        uint16_t* __restrict__ oa, uint16_t* __restrict__ od, int32_t* __restrict__ labels) {
    const int64_t jl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (jl >= M_local) return;
    const uint32_t j = (uint32_t)(cell_offset + jl);
    PhiloxSeq s(key, 0, 0, STREAM_SYNTH, j);
    double u, v;
    s.next2(u, v);
    int lab = (int)(u * (double)K);
    if (lab >= K) lab = K - 1;
    if (FILL) labels[jl] = lab + 1;
    int64_t i = -1, n = 0;
    const int64_t o0 = FILL ? p[jl] : 0;
    for (;;) {
        s.next2(u, v);
        const double gap = floor(log(u) / log1m_cov);
        if (gap >= (double)(N - i)) break;
        i += 1 + (int64_t)gap;
        if (i >= N) break;
        if (FILL) {
            const uint32_t d = c_depth_q[(uint32_t)(v * 256.0) & 255u];
            double ua, ub;
            s.next2(ua, ub);
            uint32_t a;
            if ((cfg[i] >> lab) & 1u) {
                a = binom_inv(d, theta1[i], ua);
            } else {
                PhiloxSeq t(key, 0, (uint32_t)i + 1u, STREAM_SYNTH + 1, j);
                const double th0 = beta_draw(0.2, 99.8, t);
                a = binom_inv(d, th0, ua);
            }
            oi[o0 + n] = (int32_t)i;
            oa[o0 + n] = (uint16_t)a;
            od[o0 + n] = (uint16_t)d;
        } else {
            double ua, ub;
            s.next2(ua, ub);   // keep the stream aligned with the fill pass
        }
        ++n;
    }
    if (!FILL) counts[jl] = n;
}

}  // namespace

void synth_generate(Donor& d, int64_t N, int64_t M_total, int64_t cell_offset, int64_t M_local, int K,
                    const double* Config, double coverage, uint64_t seed, cudaStream_t st) {
    if (K < 1 || K > KMAX) throw Error(5, "K must be in 1..32");
    if (!(coverage > 0.0 && coverage < 1.0)) throw Error(1, "coverage must be in (0, 1)");
    if (M_total >= (int64_t)0xFFFFFFFFll) throw Error(5, "too many cells");
    PhiloxKey key{(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    std::vector<uint32_t> cfg((size_t)N, 0u);
    std::vector<double> th1((size_t)N);
    for (int64_t i = 0; i < N; ++i) {
        for (int k = 0; k < K; ++k)
            if (Config[i + (int64_t)k * N] != 0.0) cfg[(size_t)i] |= 1u << k;
        PhiloxSeq s(key, 0, 0, STREAM_SYNTH + 2, (uint32_t)i);
        th1[(size_t)i] = beta_draw(0.45, 0.55, s);
    }
    uint32_t* cfg_d = nullptr; double* th1_d = nullptr; int64_t* cnt_d = nullptr; int64_t* p_d = nullptr;
    int32_t* oi = nullptr; uint16_t* oa = nullptr; uint16_t* od = nullptr; int32_t* labels = nullptr;
    auto cleanup = [&]() {
        cudaFree(cfg_d); cudaFree(th1_d); cudaFree(cnt_d); cudaFree(p_d); cudaFree(oi); cudaFree(oa); cudaFree(od);
    };
    try {
        CDL_CUDA(cudaMalloc(&cfg_d, 4 * (size_t)N));
        CDL_CUDA(cudaMalloc(&th1_d, 8 * (size_t)N));
        CDL_CUDA(cudaMalloc(&cnt_d, 8 * (size_t)(M_local + 1)));
        CDL_CUDA(cudaMalloc(&p_d, 8 * (size_t)(M_local + 1)));
        CDL_CUDA(cudaMalloc(&labels, 4 * (size_t)M_local));
        CDL_CUDA(cudaMemcpyAsync(cfg_d, cfg.data(), 4 * (size_t)N, cudaMemcpyHostToDevice, st));
        CDL_CUDA(cudaMemcpyAsync(th1_d, th1.data(), 8 * (size_t)N, cudaMemcpyHostToDevice, st));
        const double l1m = log(1.0 - coverage);
        const unsigned grid = (unsigned)((M_local + 127) / 128);
        k_synth<false><<<grid, 128, 0, st>>>(N, M_local, cell_offset, K, l1m, cfg_d, th1_d, key, cnt_d, nullptr,
                                             nullptr, nullptr, nullptr, nullptr);
        CDL_CUDA(cudaGetLastError());
        std::vector<int64_t> cnt((size_t)M_local + 1), p((size_t)M_local + 1);
        CDL_CUDA(cudaMemcpyAsync(cnt.data(), cnt_d, 8 * (size_t)M_local, cudaMemcpyDeviceToHost, st));
        CDL_CUDA(cudaStreamSynchronize(st));
        p[0] = 0;
        for (int64_t j = 0; j < M_local; ++j) p[(size_t)j + 1] = p[(size_t)j] + cnt[(size_t)j];
        const int64_t nnz = p[(size_t)M_local];
        CDL_CUDA(cudaMemcpyAsync(p_d, p.data(), 8 * (size_t)(M_local + 1), cudaMemcpyHostToDevice, st));
        CDL_CUDA(cudaMalloc(&oi, 4 * (size_t)std::max<int64_t>(nnz, 1)));
        CDL_CUDA(cudaMalloc(&oa, 2 * (size_t)std::max<int64_t>(nnz, 1)));
        CDL_CUDA(cudaMalloc(&od, 2 * (size_t)std::max<int64_t>(nnz, 1)));
        k_synth<true><<<grid, 128, 0, st>>>(N, M_local, cell_offset, K, l1m, cfg_d, th1_d, key, nullptr, p_d, oi, oa,
                                            od, labels);
        CDL_CUDA(cudaGetLastError());
        CDL_CUDA(cudaStreamSynchronize(st));
        cudaFree(d.labels);
        d.labels = nullptr;
        d.M_total = M_total;
        d.cell_offset = cell_offset;
        donor_build(d, N, M_local, nnz, p_d, oi, oa, od, st);
        d.labels = labels;
        d.M_total = M_total;
        d.cell_offset = cell_offset;
    } catch (...) {
        cleanup();
        cudaFree(labels);
        throw;
    }
    cleanup();
}

}  // namespace cdl
