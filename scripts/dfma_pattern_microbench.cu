// How fast does the FP64 pipe run the instruction pattern of the y convolution (k3_szmap.cu / k3w_szmap.cu phase B)?
//   acc[k] = fma(tap[j], x, acc[k])   -- 22 accumulators, 28 taps in registers, x shared by up to 22 consecutive DFMAs
// against the usual peak pattern a = fma(a, m, c) (both multiplicands constant: register reuse cache).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_pattern scripts/dfma_pattern_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NB = 28, UB = 22;

template <bool SMEM>
__global__ void __launch_bounds__(512, 1) conv_pattern(double* out, const double* in, int reps, int H) {
    __shared__ double xs[96 * 33];
    for (int i = threadIdx.x; i < 96 * 33; i += blockDim.x) xs[i] = in[i % 1024];
    __syncthreads();
    double tap[NB], acc[UB];
#pragma unroll
    for (int j = 0; j < NB; ++j) tap[j] = in[j + (threadIdx.x & 31)];
#pragma unroll
    for (int k = 0; k < UB; ++k) acc[k] = 0.0;
    const int lane = threadIdx.x & 31;
    const int u0 = ((threadIdx.x >> 5) & 3) * UB;
    double xr = in[threadIdx.x];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int ii = 0; ii < UB + 2 * (NB - 1); ++ii) {
            double x;
            if (SMEM) {
                const int up = u0 - (NB - 1) + ii, ua = up < 0 ? -up : up;
                x = ua < H ? xs[ua * 33 + lane] : 0.0;
            } else {
                x = xr; xr = __longlong_as_double(__double_as_longlong(xr) ^ ii);
            }
#pragma unroll
            for (int k = 0; k < UB; ++k) {
                const int j = ii - (NB - 1) - k < 0 ? k + (NB - 1) - ii : ii - (NB - 1) - k;
                if (j < NB) acc[k] = fma(tap[j], x, acc[k]);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < UB; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512, 1) peak_pattern(double* out, int iters) {
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

// two fresh 64-bit sources per DFMA, no table structure: acc[k] = fma(t[k], x, acc[k]) with t[] as wide as acc[]
__global__ void __launch_bounds__(512, 1) fresh_pattern(double* out, const double* in, int iters) {
    double a[16], t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = 0.0; t[i] = in[i + (threadIdx.x & 31)]; }
    double x = in[threadIdx.x];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(t[k], x, a[k]);
        x = __longlong_as_double(__double_as_longlong(x) ^ i);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out, *in;
    cudaMalloc(&out, sizeof(double) * sms * 512);
    cudaMalloc(&in, sizeof(double) * 2048);
    double h[2048];
    for (int i = 0; i < 2048; ++i) h[i] = 1.0 / (1 + i);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double clk = prop.clockRate * 1e3;
    auto report = [&](const char* name, double fma_per_thread, int threads, float ms) {
        const double per_clk_smsp = fma_per_thread * threads / 32.0 / 4.0 / (ms * 1e-3 * clk);
        printf("%-44s threads/SM %3d  %8.3f ms  DFMA warp-instr / clk / SMSP = %.3f (peak 0.5)  %.1f TFLOP/s\n", name, threads, ms,
               per_clk_smsp, 2.0 * fma_per_thread * threads * sms / (ms * 1e-3) / 1e12);
    };
    for (int threads : {512, 256, 128}) {
        float ms;
        const int reps = 200;
        const double fmas = 1210.0 * reps;
        for (int w = 0; w < 2; ++w) {
            cudaEventRecord(e0); conv_pattern<true><<<sms, threads>>>(out, in, reps, 86); cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1); report("conv pattern, x from shared memory", fmas, threads, ms);
        for (int w = 0; w < 2; ++w) {
            cudaEventRecord(e0); conv_pattern<false><<<sms, threads>>>(out, in, reps, 86); cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1); report("conv pattern, x from a register", fmas, threads, ms);
        const int iters = 20000;
        for (int w = 0; w < 2; ++w) {
            cudaEventRecord(e0); fresh_pattern<<<sms, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1); report("a[k] = fma(t[k], x, a[k]) (two fresh sources)", 16.0 * iters, threads, ms);
        for (int w = 0; w < 2; ++w) {
            cudaEventRecord(e0); peak_pattern<<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1); report("a[k] = fma(a[k], m, c) (peak pattern)", 16.0 * iters, threads, ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
