// Micro-benchmark (tuning aid): float64 instruction cost on one SM -- dependent-chain latency (1 warp) and throughput (32 warps,
// 4 independent chains each) of DFMA, DADD, DSETP+select and F2F.F64.F32, beside FFMA for scale.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f64_rate f64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(int iters, long long *cyc, double *sink, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
    float f0 = (float)a0, f1 = (float)a1, f2 = (float)a2, f3 = (float)a3;
    const double m = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c); }
        if (OP == 1) { a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c; }
        if (OP == 2) { a0 = a0 > a1 ? a0 + c : a1; a2 = a2 > a3 ? a2 + c : a3; a1 = a1 > a2 ? a1 : a2 + c; a3 = a3 > a0 ? a3 : a0 + c; }
        if (OP == 3) { a0 += (double)f0; f0 = __int_as_float(__float_as_int(f0) ^ i); a1 += (double)f1; f1 = __int_as_float(__float_as_int(f1) ^ i);
                       a2 += (double)f2; f2 = __int_as_float(__float_as_int(f2) ^ i); a3 += (double)f3; f3 = __int_as_float(__float_as_int(f3) ^ i); }
        if (OP == 4) { f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f); }
        if (OP == 5) { a0 = fma(a0, m, c); }                       // one dependent chain: latency
        if (OP == 6) { a0 = a0 + c; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3 == 12345.678) sink[0] = a0;
}

int main() {
    long long *d, h;
    double *sink;
    cudaMalloc(&d, 8); cudaMalloc(&sink, 8);
    const int iters = 20000;
    const char *names[] = {"DFMA x4 chains", "DADD x4 chains", "DSETP+sel+DADD x4", "F2F.F64.F32+DADD x4", "FFMA x4 chains", "DFMA 1 chain", "DADD 1 chain"};
    for (int op = 0; op < 7; op++)
        for (int threads : {32, 1024}) {
            for (int rep = 0; rep < 2; rep++) {
                switch (op) {
                case 0: k<0><<<1, threads>>>(iters, d, sink, 1.0); break;
                case 1: k<1><<<1, threads>>>(iters, d, sink, 1.0); break;
                case 2: k<2><<<1, threads>>>(iters, d, sink, 1.0); break;
                case 3: k<3><<<1, threads>>>(iters, d, sink, 1.0); break;
                case 4: k<4><<<1, threads>>>(iters, d, sink, 1.0); break;
                case 5: k<5><<<1, threads>>>(iters, d, sink, 1.0); break;
                case 6: k<6><<<1, threads>>>(iters, d, sink, 1.0); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const int per_iter = op >= 5 ? 1 : 4;
            printf("%-22s %4d threads: %.2f cycles per loop iteration (%d ops per thread) -> %.2f warp-instructions per cycle per SM\n", names[op], threads,
                   (double)h / iters, per_iter, (double)per_iter * (threads / 32) * iters / (double)h);
        }
    return 0;
}
