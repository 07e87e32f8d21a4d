// Measures fp64 vector-pipe throughput (DFMA / DADD / DMUL) per SM on the device it runs on, and what a
// PREDICATED DADD costs when its predicate is false (op 3: all lanes off, op 4: bit pattern 0x1111 -> one
// clone in four on), which decides how the masked accumulate of k_cells_sell should be written.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double x, double y, uint32_t bits) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = fma(a[i], y, x);
            else if (OP == 1) a[i] = a[i] + y;
            else if (OP == 2) a[i] = a[i] * y;
            else {
                asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.eq.u32 p, t, 0;\n\t"
                             "@p bra SKIP%=;\n\tadd.f64 %0, %0, %1;\n\tSKIP%=:\n\t}"
                             : "+d"(a[i]) : "d"(y), "r"(bits), "r"(1u << i));
            }
        }
        if (OP >= 3) bits = (bits << 8) | (bits >> 24);   // keep the predicates loop-variant
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
    const char* names[] = {"DFMA", "DADD", "DMUL", "@P DADD all-off", "@P DADD 1-in-4 on", "@P DADD all-on"};
    for (int op = 0; op < 6; ++op) for (int threads : {256, 512}) for (int bps : {1, 2}) {
        if (threads * bps > 2048) continue;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k<0><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999, 0);
            if (op == 1) k<1><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999, 0);
            if (op == 2) k<2><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999, 0);
            if (op == 3) k<3><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999, 0u);
            if (op == 4) k<3><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999, 0x11111111u);
            if (op == 5) k<3><<<sms * bps, threads>>>(out, iters, 1.0000001, 0.9999999, 0xffffffffu);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * bps * threads * iters * 8;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("op=%s threads=%d blocks/SM=%d: %.2f Tinst/s  = %.1f lane-ops/clk/SM (at %d MHz nominal)\n",
               names[op], threads, bps, ops / ms / 1e9, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    }
    return 0;
}
