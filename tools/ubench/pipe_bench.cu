// Micro-benchmark (tuning aid): issue throughput per SM sub-partition of the instructions the conversion / recursion code is
// made of.  16 warps per SM (4 per scheduler), 8 independent dependency chains per thread, all operands in registers.
// Prints cycles per warp-instruction per scheduler (1.0 = full rate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, long long *cyc, uint32_t seed) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = 0x3f800000u + ((seed + threadIdx.x * 16 + i) & 0xffff);   // floats in [1, 1.008)
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int a = 2 * i, b = 2 * i + 1;
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float *)&r[a]) : "f"(*(float *)&r[b]));
            if (OP == 1) asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%0, %1};\n\tfma.rn.f32x2 t, t, t, t;\n\tmov.b64 {%0, %1}, t;\n\t}" : "+r"(r[a]), "+r"(r[b]));
            if (OP == 2) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %0;" : "+r"(r[a]) : "f"(*(float *)&r[b]));
            if (OP == 3) asm volatile("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %1;\n\tfma.rn.f32.f16 %0, h0, h1, %0;\n\t}" : "+f"(*(float *)&r[a]) : "r"(r[b]));
            if (OP == 4) asm volatile("{\n\t.reg .b16 h0, h1;\n\tmov.b32 {h0, h1}, %0;\n\tcvt.f32.f16 %0, h0;\n\t}" : "+r"(r[a]));
            if (OP == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(*(float *)&r[a]));
            if (OP == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float *)&r[a]) : "f"(*(float *)&r[b]));
            if (OP == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(r[a]) : "r"(r[b]));
            if (OP == 8) asm volatile("add.f32 %0, %0, %1;" : "+f"(*(float *)&r[a]) : "f"(*(float *)&r[b]));
            if (OP == 9) asm volatile("{\n\t.reg .b64 t, s;\n\tmov.b64 t, {%0, %1};\n\tadd.f32x2 t, t, t;\n\tmov.b64 {%0, %1}, t;\n\t}" : "+r"(r[a]), "+r"(r[b]));
            if (OP == 10) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(r[a]) : "r"(r[b]));
        }
    }
    long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, uint32_t *out, long long *cyc) {
    for (int rep = 0; rep < 2; rep++) { k<OP><<<148, 512>>>(out, cyc, 1u); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double per_sched = 4.0 * 8 * ITERS;      // warp instructions of the measured kind per scheduler
    printf("%-30s %6.2f cycles per warp-instruction per scheduler\n", name, avg / per_sched);
}
int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA", out, cyc); run<8>("FADD", out, cyc); run<1>("FFMA2 (fma.f32x2)", out, cyc); run<9>("FADD2 (add.f32x2)", out, cyc);
    run<6>("FMNMX", out, cyc); run<7>("SHF", out, cyc); run<10>("LOP3", out, cyc);
    run<2>("F2FP.F16.F32.PACK_AB", out, cyc); run<3>("FHFMA (f32 = f16*f16 + f32)", out, cyc); run<4>("HADD2.F32 (f16 -> f32)", out, cyc);
    run<5>("MUFU.EX2", out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    return 0;
}
