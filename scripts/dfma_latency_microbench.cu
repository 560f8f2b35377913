// DFMA on B200: dependent-issue latency and throughput against the number of independent chains per thread and of warps
// per scheduler.  Pattern of the y convolution: a[k] = fma(t[k], x, a[k]) with x changing every pass over the chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dfma_latency.bin scripts/dfma_latency_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chains(double* out, const double* in, int iters, long long* cyc) {
    double a[C], t[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { a[k] = in[threadIdx.x + k]; t[k] = in[threadIdx.x + 64 + k]; }
    double x = in[threadIdx.x + 200];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < C; ++k) a[k] = fma(t[k], x, a[k]);
            x += 1e-9;                       // a new x per pass (one DADD per C DFMA)
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < C; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int C>
void run(double* out, double* in, long long* cyc) {
    const int iters = 2000;
    for (int warps_per_sm : {4, 8, 16}) {
        chains<C><<<148, 32 * warps_per_sm>>>(out, in, iters, cyc);
        chains<C><<<148, 32 * warps_per_sm>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        const double dfma_per_warp = (double)iters * 8 * C;
        const double per_clk_sched = dfma_per_warp * (warps_per_sm / 4.0) / (double)h;
        printf("chains %2d  warps/scheduler %d  cycles per DFMA per warp %6.2f   DFMA / clk / scheduler %.3f\n", C, warps_per_sm / 4,
               (double)h / dfma_per_warp, per_clk_sched);
    }
}

int main() {
    double *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&in, 4096 * 8); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096 * 8);
    run<1>(out, in, cyc); run<2>(out, in, cyc); run<4>(out, in, cyc); run<8>(out, in, cyc);
    run<16>(out, in, cyc); run<22>(out, in, cyc); run<32>(out, in, cyc);
    printf("rc=%d\n", (int)cudaGetLastError());
    return 0;
}
