// Developer microbenchmark: throughput of the legacy warp-level INT8 tensor instruction on B200
// (mma.sync.m16n8k32.s8.s8.s32), to decide whether an Ozaki-split (int8 slices, exact int32 accumulation)
// of the FP64 GEMMs could beat the DMMA path without going to tcgen05.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_peak scripts/imma_peak_microbench.cu && ./imma_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void imma_kernel(int* out, int iters) {
    int c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
    unsigned a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x01020304u, b1 = 0x04030201u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 1 << 14;
    int* buf;
    cudaMalloc(&buf, sizeof(int) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    imma_kernel<<<blocks, threads>>>(buf, iters);
    double best = 0;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        imma_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tops = 2.0 * 16 * 8 * 32 * 8.0 * iters * blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (tops > best) best = tops;
    }
    printf("%s: mma.sync m16n8k32 s8: %.1f TOPS (dense int8; tcgen05 nominal 4500)\n", p.name, best);
    return cudaGetLastError() != cudaSuccess;
}
