// Measures fp64 vector-pipe throughput (DFMA / DADD) per SM on the device it runs on.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double x, double y) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fma(a[i], y, x);
            else if (OP == 1) a[i] = a[i] + y;
            else a[i] = a[i] * y;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int op = 0; op < 3; ++op) for (int threads : {128, 256, 512}) for (int bps : {1, 2, 4}) {
        if (threads * bps > 2048) continue;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k<0><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999);
            if (op == 1) k<1><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999);
            if (op == 2) k<2><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * bps * threads * iters * 8;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("op=%s threads=%d blocks/SM=%d: %.2f Tinst/s  = %.1f lane-ops/clk/SM (at %d MHz nominal)\n",
               op == 0 ? "DFMA" : op == 1 ? "DADD" : "DMUL", threads, bps, ops / ms / 1e9,
               ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
    return 0;
}
