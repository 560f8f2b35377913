// Peak-throughput microbenchmarks used by bench.py for the roofline denominators that MEASURED_PEAKS.json does
// not carry (FP64 FMA and FP64 tensor-core DMMA rates).  Measurement tooling only: built into
// scripts/libjx_peaks.so, never linked into the product library libjoxsz_b200.so.
#include <cuda_runtime.h>
#include <stdint.h>

#define JX_OK 0
#define JX_ERR_INVALID -1
#define JX_ERR_CUDA -2

namespace {
__global__ void fp64_peak_kernel(double* out, int iters) {
    // 16 independent chains per thread, both multiplicands constant (register reuse cache): the pattern that reaches the
    // highest DFMA rate on B200 (36.3 TFLOP/s; 8 chains give 34.8) -- see scripts/dfma_pattern_microbench.cu
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_peak_kernel(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int jxp_measure_dmma_tflops(int32_t device, double* tflops) {
    if (!tflops) return JX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return JX_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 1 << 12;
    double* buf = nullptr;
    if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return JX_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dmma_peak_kernel<<<blocks, threads>>>(buf, iters);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
        double tf = 512.0 * 8.0 * (double)iters * blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaError_t e = cudaGetLastError();
    *tflops = best;
    return e == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}

extern "C" int jxp_measure_fp64_tflops(int32_t device, double* tflops) {
    if (!tflops) return JX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return JX_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double* buf = nullptr;
    if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return JX_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads>>>(buf, iters);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 16.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaError_t e = cudaGetLastError();
    *tflops = best;
    return e == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}
