// Which instructions share the fp64 pipe of an sm_100 SM?  Times mixes of the instructions of k_cells_sell's inner
// loop (cells_sell.cu): predicated DADDs fed by R2P, I2F.F64.U16 count conversions, and the 2^52 "magic number"
// conversion that replaces them with one integer instruction + an exact DFMA.  Every variant runs the same
// 16 predicated adds per "entry"; what differs is how the two counts become doubles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu && ./fp64_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void madd16(double (&acc)[16], double del, uint32_t m) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int bp = k < 7 ? k : (k < 14 ? 8 + (k - 7) : (k == 14 ? 7 : 15));
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.eq.u32 p, t, 0;\n\t"
                     "@p bra SKIP%=;\n\tadd.f64 %0, %0, %1;\n\tSKIP%=:\n\t}"
                     : "+d"(acc[k]) : "d"(del), "r"(m), "r"(1u << bp));
    }
}

// MODE 0: del = du (no conversion at all: the floor)        MODE 1: two I2F.F64.U16 + DMUL + DFMA (round 1 kernel)
// MODE 2: magic-number conversion: (2^52 + a) built with integer ops, a*du = fma(2^52 + a, du, -2^52 du) exactly
//         rounded, likewise b*dv, one DADD: three fp64 instructions, no I2F
// MODE 3: I2F.F64.U16 only (its own rate)                    MODE 4: plain DADD only (16 per entry)
template <int MODE>
__global__ void k(double* out, const uint32_t* in, int iters, double du0, double dv0) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.0;
    uint32_t w = in[threadIdx.x & 31], m = in[32 + (threadIdx.x & 31)];
    double du = du0 + threadIdx.x * 1e-9, dv = dv0;
    for (int it = 0; it < iters; ++it) {
        double del;
        if (MODE == 0) del = du;
        if (MODE == 1 || MODE == 3) {
            double a, b;
            asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.rn.f64.u16 %0, l;\n\t}" : "=d"(a) : "r"(w));
            asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.rn.f64.u16 %0, h;\n\t}" : "=d"(b) : "r"(w));
            del = fma(a, du, b * dv);
        }
        if (MODE == 2) {
            const double A = __hiloint2double(0x43300000, (int)(w & 0xffffu));
            const double B = __hiloint2double(0x43300000, (int)(w >> 16));
            const double c1 = __hiloint2double(__double2hiint(du) + (int)0x83400000, __double2loint(du));   // -2^52 du
            const double c2 = __hiloint2double(__double2hiint(dv) + (int)0x83400000, __double2loint(dv));
            del = fma(A, du, c1) + fma(B, dv, c2);
        }
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += du;
        } else if (MODE == 3) {
            acc[it & 1] += del;          // (kept cheap: one DADD per two conversions)
        } else {
            madd16(acc, del, m);
        }
        w = (w >> 3) | (w << 29);
        m = (m >> 5) | (m << 27);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 896, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * sms * threads);
    uint32_t h[64]; for (int i = 0; i < 64; ++i) h[i] = 0x9e3779b9u * (i + 1);
    uint32_t* in; cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[] = {"16 @P DADD, no conversion", "16 @P DADD + 2 I2F.F64.U16 + DMUL + DFMA",
                           "16 @P DADD + magic-number conversion (2 DFMA + DADD, integer ops)",
                           "2 I2F.F64.U16 + DMUL + DFMA + 1 DADD", "16 plain DADD"};
    for (int mode = 0; mode < 5; ++mode) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<sms, threads>>>(out, in, iters, 0.25, -0.5);
            if (mode == 1) k<1><<<sms, threads>>>(out, in, iters, 0.25, -0.5);
            if (mode == 2) k<2><<<sms, threads>>>(out, in, iters, 0.25, -0.5);
            if (mode == 3) k<3><<<sms, threads>>>(out, in, iters, 0.25, -0.5);
            if (mode == 4) k<4><<<sms, threads>>>(out, in, iters, 0.25, -0.5);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double cyc = ms * 1e-3 * clk * 1e3;                       // SM cycles at the nominal clock
        const double per_entry = cyc / ((double)iters * threads / 32 / 4);   // cycles per warp-entry per SM sub-partition
        printf("mode %d (%s): %.3f ms, %.1f cycles per warp-entry per sub-partition (28 warps/SM, %d MHz nominal)\n",
               mode, names[mode], ms, per_entry, clk / 1000);
    }
    return 0;
}
