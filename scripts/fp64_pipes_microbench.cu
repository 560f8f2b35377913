// Are the FP64 FMA pipe and the FP64 tensor (DMMA) pipe of B200 separate?  Run DFMA-only, DMMA-only and a mix
// (even warps DFMA, odd warps DMMA) and compare the per-kind throughputs.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mix(double* out, int iters, int mode) {
    const int warp = threadIdx.x >> 5;
    const bool do_dmma = mode == 1 || (mode == 2 && (warp & 1));
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
    const double a = 1.0000001, b = 1e-7;
    if (do_dmma) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    } else {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 2, threads = 512, iters = 1 << 13;
    double* buf; cudaMalloc(&buf, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        mix<<<blocks, threads>>>(buf, iters, mode);
        cudaEventRecord(e0); mix<<<blocks, threads>>>(buf, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double warps = (double)blocks * threads / 32;
        double dfma_w = mode == 0 ? warps : mode == 1 ? 0 : warps / 2, dmma_w = warps - dfma_w;
        double tf_dfma = dfma_w * 32 * 16.0 * 2 * iters / (ms * 1e-3) / 1e12;   // 16 fma per thread per iter
        double tf_dmma = dmma_w * 8 * 512.0 * iters / (ms * 1e-3) / 1e12;
        printf("mode %d: %.3f ms  DFMA %.1f TF/s  DMMA %.1f TF/s  sum %.1f\n", mode, ms, tf_dfma, tf_dmma, tf_dfma + tf_dmma);
    }
    return 0;
}
