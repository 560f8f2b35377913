// The inner loop of the large-map y convolution (k3l2_szmap.cu: k3m_yconv on a padded shared-memory tile) in
// isolation: 256 threads per SM, one 16-row block per warp and pass.  Variants: stores on / off, x from shared memory
// or from a register, tasks back to back or separated by __syncthreads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/yconv_loop.bin scripts/yconv_loop_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NB = 28, UB = 16, PAD = NB - 1, H = 128, ROWS = H + 2 * PAD + NB;

template <int FROM_SMEM>
__device__ __forceinline__ void yconv(const double* tile, int lane, int u0, const double (&tap)[NB], double (&acc)[UB], double xr) {
    constexpr int NIN = UB + 2 * PAD;
    const double* in = tile + u0 * 32 + lane;
#pragma unroll
    for (int k = 0; k < UB; ++k) acc[k] = 0.0;
    double xa[FROM_SMEM >= 2 ? NIN : 1];
    if (FROM_SMEM >= 2) {                       // every input of the block in registers before the first DFMA
#pragma unroll
        for (int ii = 0; ii < NIN; ++ii) xa[ii] = in[ii * 32];
        if (FROM_SMEM == 3) asm volatile("" ::: "memory");
    }
#pragma unroll
    for (int ii = 0; ii < NIN; ++ii) {
        double x;
        if (FROM_SMEM >= 2) x = xa[ii]; else if (FROM_SMEM) x = in[ii * 32]; else { xr += 1e-9; x = xr; }
#pragma unroll
        for (int k = 0; k < UB; ++k) {
            const int j = ii - PAD - k < 0 ? k + PAD - ii : ii - PAD - k;
            if (j < NB) acc[k] = fma(tap[j], x, acc[k]);
        }
    }
}

template <int FROM_SMEM, int STORES, int SYNC, int SKEW, int NT = 256>
__global__ void __launch_bounds__(NT, 1) k(double* out, const double* in, int iters, long long* cyc) {
    extern __shared__ double tile[];
    for (int i = threadIdx.x; i < ROWS * 32; i += NT) tile[i] = in[i % 4096];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double tap[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) tap[j] = tile[(H + 2 * PAD + j) * 32 + lane];
    double* o = out + ((size_t)blockIdx.x * NT + threadIdx.x) * UB;
    double sum = 0.0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (STORES == 7) {
            // nine warps: warps 0-7 compute and stage their results in shared memory, warp 8 does all the global traffic
            // (the tile fetch by cp.async and the flush of the staged results by plain coalesced stores)
            if (warp == 8) {
                __syncthreads();               // X: (nothing to do for the memory warp)
                __syncthreads();               // Y: the staged results are complete
                double* g = out + (size_t)blockIdx.x * 4096;
                const double* st = tile + ROWS * 32;
#pragma unroll 8
                for (int q = 0; q < 128; ++q) g[q * 32 + lane] = st[q * 32 + lane];
#pragma unroll
                for (int q = 0; q < 96; ++q)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(
                                     tile + ROWS * 32 + 4096 + 2 * (q * 32 + lane))), "l"(in + 2 * ((q * 32 + lane) % 2048)) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                double acc[UB];
                yconv<FROM_SMEM>(tile, lane, 16 * warp, tap, acc, sum);
                __syncthreads();               // X: the memory warp has flushed the previous round
#pragma unroll
                for (int kk = 0; kk < UB; ++kk) tile[ROWS * 32 + (16 * warp + kk) * 32 + lane] = acc[kk];
                __syncthreads();               // Y
            }
            continue;
        }
        if (SKEW == -1) {              // the kernel's tile fetch: global -> shared copies by the computing warps themselves
#pragma unroll
            for (int q = 0; q < 12; ++q)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(
                                 tile + ROWS * 32 + UB * NT + 2 * (q * NT + threadIdx.x))), "l"(in + 2 * ((q * NT + threadIdx.x) % 2048)) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        double acc[UB];
        yconv<FROM_SMEM>(tile, lane, 16 * (warp & 7), tap, acc, sum);
        if (STORES == 1) {
#pragma unroll
            for (int kk = 0; kk < UB; ++kk) __stcg(o + kk, acc[kk]);
        } else if (STORES == 2) {
#pragma unroll
            for (int kk = 0; kk < UB; ++kk) o[kk] = acc[kk];
        } else if (STORES == 3) {
#pragma unroll
            for (int kk = 0; kk < UB; ++kk) tile[ROWS * 32 + kk * NT + threadIdx.x] = acc[kk];
        } else if (STORES == 6) {
            // results staged in shared memory, flushed by ONE bulk copy (async proxy) per task round
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < UB; ++kk) tile[ROWS * 32 + (16 * (warp & 7) + kk) * 32 + lane] = acc[kk];
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (size_t)blockIdx.x * 4096),
                             "r"((unsigned)__cvta_generic_to_shared(tile + ROWS * 32)), "r"(32768) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (STORES == 4 || STORES == 5) {
            // deferred: see below
        } else {
#pragma unroll
            for (int kk = 0; kk < UB; ++kk) asm volatile("" ::"d"(acc[kk]));
        }
        if (SYNC && (it % SYNC) == SYNC - 1) {
            if (STORES == 5) asm volatile("bar.sync 1, %0;" ::"n"(NT)); else __syncthreads();
            if (SKEW > 0 && warp >= 4) __nanosleep(SKEW);
        }
        if (STORES == 4) {
#pragma unroll
            for (int kk = 0; kk < UB; ++kk) __stcg(o + kk, acc[kk]);
        }
        if (STORES == 5) {
#pragma unroll
            for (int kk = 0; kk < UB; ++kk) __stcg(o + kk, acc[kk]);
        }
    }
    const long long t1 = clock64();
    if (STORES == 6 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (STORES == 0) o[0] = sum;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int A, int B, int C, int D, int NT = 256>
void run(const char* name, double* out, double* in, long long* cyc) {
    const int iters = 400;
    const size_t smem = (ROWS * 32 + UB * NT + 24 * NT + 8192) * sizeof(double);
    cudaFuncSetAttribute(k<A, B, C, D, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<A, B, C, D, NT><<<148, NT, smem>>>(out, in, iters, cyc);
    k<A, B, C, D, NT><<<148, NT, smem>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const double dfma = (double)iters * UB * (2 * NB - 1);
    printf("%-58s cycles per task %7.0f   DFMA / clk / scheduler %.3f  (%s)\n", name, (double)h / iters, ((NT == 288 ? 256 : NT) / 128.0) * dfma / (double)h,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double *out, *in; long long* cyc;
    cudaMalloc(&out, (size_t)148 * 512 * UB * 8); cudaMalloc(&in, 4096 * 8); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096 * 8);
    run<1, 1, 1, 0>("smem x, stores, barrier per task (as shipped)", out, in, cyc);
    run<1, 1, 2, 0>("smem x, stores, barrier every 2 tasks", out, in, cyc);
    run<1, 1, 4, 0>("smem x, stores, barrier every 4 tasks", out, in, cyc);
    run<1, 1, 0, 0>("smem x, stores, no barrier", out, in, cyc);
    run<1, 1, 1, 200>("smem x, stores, barrier per task, warps 4-7 sleep 200 ns", out, in, cyc);
    run<1, 1, 1, 1000>("smem x, stores, barrier per task, warps 4-7 sleep 1 us", out, in, cyc);
    run<0, 1, 1, 0>("register x, stores, barrier per task", out, in, cyc);
    run<0, 1, 0, 0>("register x, stores, no barrier", out, in, cyc);
    run<2, 1, 1, 0>("smem x preloaded (source order), stores, barrier per task", out, in, cyc);
    run<3, 1, 1, 0>("smem x preloaded + compiler barrier, stores, barrier per task", out, in, cyc);
    run<1, 7, 0, 0, 288>("nine warps: eight compute + stage, one memory warp (flush by plain stores + cp.async fetch)", out, in, cyc);
    run<1, 3, 1, -1>("smem x, shared-memory stores, barrier per task, + 12 cp.async per thread and task", out, in, cyc);
    run<1, 3, 1, 0>("smem x, shared-memory stores, barrier per task (again)", out, in, cyc);
    run<1, 6, 0, -1>("smem x, staged + bulk store, + 12 cp.async per thread and task", out, in, cyc);
    run<1, 6, 0, 0>("smem x, results staged + one 32 KB bulk store per round (2 barriers)", out, in, cyc);
    run<1, 4, 1, 0>("smem x, stores AFTER the barrier of their task", out, in, cyc);
    run<1, 5, 1, 0>("smem x, stores after a named barrier without memory clobber", out, in, cyc);
    run<1, 2, 1, 0>("smem x, plain global stores, barrier per task", out, in, cyc);
    run<1, 3, 1, 0>("smem x, shared-memory stores, barrier per task", out, in, cyc);
    run<1, 0, 1, 0>("smem x, no stores (results kept alive), barrier per task", out, in, cyc);
    run<1, 0, 0, 0>("smem x, no stores, no barrier", out, in, cyc);
    run<1, 1, 1, 0, 128>("128 threads (1 warp per scheduler): smem x, barrier per task", out, in, cyc);
    run<1, 1, 0, 0, 128>("128 threads: smem x, no barrier", out, in, cyc);
    run<0, 1, 1, 0, 128>("128 threads: register x, barrier per task", out, in, cyc);
    run<1, 1, 1, 0, 512>("512 threads (4 warps per scheduler): smem x, barrier per task", out, in, cyc);
    run<1, 1, 0, 0, 512>("512 threads: smem x, no barrier", out, in, cyc);
    return 0;
}
