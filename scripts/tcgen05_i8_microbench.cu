// What does tcgen05.mma.kind::i8 deliver on B200 when it is driven by hand (shared-memory descriptors written by the
// kernel itself, accumulators in tensor memory, tcgen05.commit -> mbarrier, tcgen05.ld for the read-back)?
//
// This grounds the cost model of the int8-slice (Ozaki) form of the FP64 GEMMs K2 / K7: the
// products of that scheme are exactly these instructions.  Part 1 checks one 128 x N x 64 product against the host
// (descriptor layout, instruction descriptor and accumulator read-back are right); part 2 measures the issue rate of
// back-to-back accumulating MMAs on resident operands -- the upper bound any such GEMM could reach.
//
// Operand layout (K-major, no swizzle): 8-row x 16-byte core matrices of 128 contiguous bytes; the two K halves of a
// 32-byte MMA step sit LBO = 128 B apart, 8-row groups SBO = 256 B x (K / 32) apart ... here one tile per MMA step.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_i8 scripts/tcgen05_i8_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: start address, leading / stride byte offsets in units of
// 16 bytes, version 1 for sm_100, layout type 0 = no swizzle)
__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                       // version_
    return d;
}

// instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(n));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// byte offset of element (row, k) of an operand tile with `ksteps` MMA steps of 32 bytes: step-major, then 8-row group
// (256 B), then K half (128 B), then row in group (16 B)
__host__ __device__ inline size_t tile_off(int row, int k, int rows) {
    const int step = k >> 5, kk = k & 31;
    return (size_t)step * rows * 32 + (size_t)(row >> 3) * 256 + (size_t)(kk >> 4) * 128 + (size_t)(row & 7) * 16 + (kk & 15);
}

template <int N>
__global__ void __launch_bounds__(128, 1) i8_kernel(const int8_t* __restrict__ a_g, const int8_t* __restrict__ b_g, int ksteps,
                                                     int reps, int32_t* __restrict__ d_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    int8_t* a_s = reinterpret_cast<int8_t*>(smem);                    // [ksteps][128 x 32]
    int8_t* b_s = a_s + (size_t)ksteps * 128 * 32;                    // [ksteps][N x 32]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < ksteps * 128 * 32; i += 128) a_s[i] = a_g[i];
    for (int i = tid; i < ksteps * N * 32; i += 128) b_s[i] = b_g[i];
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");    // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_slot;
    constexpr uint32_t idesc = make_idesc(N);
    if (tid == 0) {
        for (int r = 0; r < reps; ++r)
            for (int s = 0; s < ksteps; ++s)
                mma_i8(tm, make_desc(a_s + (size_t)s * 128 * 32, 128, 256), make_desc(b_s + (size_t)s * N * 32, 128, 256), idesc,
                       (r | s) ? 1u : 0u);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (d_out && blockIdx.x == 0) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + c0, r);
            for (int j = 0; j < 32; ++j) d_out[(size_t)(32 * warp + (tid & 31)) * N + c0 + j] = (int32_t)r[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tm), "r"(512u) : "memory");
}

template <int N>
int run(int sms, double clk_hz) {
    const int ksteps = 2;
    std::vector<int8_t> a((size_t)ksteps * 128 * 32), b((size_t)ksteps * N * 32);
    std::vector<int8_t> a_rm((size_t)128 * 64), b_rm((size_t)N * 64);
    srand(7 + N);
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) { int8_t v = (int8_t)(rand() % 127 - 63); a_rm[m * 64 + k] = v; a[tile_off(m, k, 128)] = v; }
    for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) { int8_t v = (int8_t)(rand() % 127 - 63); b_rm[n * 64 + k] = v; b[tile_off(n, k, N)] = v; }
    int8_t *ad, *bd; int32_t* dd;
    CK(cudaMalloc(&ad, a.size())); CK(cudaMalloc(&bd, b.size())); CK(cudaMalloc(&dd, sizeof(int32_t) * 128 * N));
    CK(cudaMemcpy(ad, a.data(), a.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(bd, b.data(), b.size(), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)ksteps * (128 + N) * 32;
    CK(cudaFuncSetAttribute(i8_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    i8_kernel<N><<<1, 128, smem>>>(ad, bd, ksteps, 1, dd);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> d(128 * N);
    CK(cudaMemcpy(d.data(), dd, sizeof(int32_t) * 128 * N, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        int32_t ref = 0;
        for (int k = 0; k < 64; ++k) ref += (int32_t)a_rm[m * 64 + k] * (int32_t)b_rm[n * 64 + k];
        if (ref != d[m * N + n]) { if (bad < 3) printf("  mismatch N=%d (%d,%d): got %d want %d\n", N, m, n, d[m * N + n], ref); ++bad; }
    }
    printf("N=%3d  correctness: %ld of %d accumulators differ from the host product (K = 64)\n", N, bad, 128 * N);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20000;
    i8_kernel<N><<<sms, 128, smem>>>(ad, bd, ksteps, 200, nullptr);
    cudaEventRecord(e0);
    i8_kernel<N><<<sms, 128, smem>>>(ad, bd, ksteps, reps, nullptr);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double macs = (double)sms * reps * ksteps * 128.0 * N * 32.0;
    printf("N=%3d  issue rate: %.3f ms for %d MMAs per SM -> %.1f TOPS (dense int8, 2 ops per MAC), %.1f cycles per 128x%dx32 MMA\n", N, ms,
           reps * ksteps, 2.0 * macs / (ms * 1e-3) / 1e12, ms * 1e-3 * clk_hz / (reps * ksteps), N);
    cudaFree(ad); cudaFree(bd); cudaFree(dd);
    return bad ? 2 : 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount; const double clk = prop.clockRate * 1e3;
    printf("%s, %d SMs, %.0f MHz\n", prop.name, sms, clk / 1e6);
    int rc = 0;
    rc |= run<64>(sms, clk);
    rc |= run<128>(sms, clk);
    rc |= run<256>(sms, clk);
    return rc;
}
