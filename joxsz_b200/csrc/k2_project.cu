// K2 -- line-of-sight projection as a dense FP64 contraction on the tensor cores.
//
//   out[W, nout] = pp[W, K] . op[nout, K]^T
//
// `op` is a constant operator built once on the host (joxsz_b200/operators.py): for the map stage it
// is (not-a-knot spline fit) o (Compton-y scaling) o (PyAbel forward direct transform), i.e. it replaces,
// for every walker at once, reference joxsz_funcs.py:457-460; with `y_op` it yields the Compton-y
// profile itself (:457-459).  The reference rebuilds four Nr x Nr matrices per call for this.
//
// The accuracy budget of the likelihood (1e-6 absolute on chi^2 ~ 1e3) needs float64 operands and
// accumulation, so the contraction runs on the FP64 tensor-core path: mma.sync.m8n8k4.f64 (DMMA).
// tcgen05 has no f64 kind; an Ozaki-split on tcgen05 is the documented alternative.
//
// Tiling: CTA 128 (walkers) x 64 (outputs), 8 warps as 4 x 2, warp tile 32 x 32 = 4 x 4 DMMA tiles.
// K is consumed in chunks of 16 through a 3-stage cp.async ring (92 KB: TWO CTAs per SM, so that one computes while
// the other fills its ring or stores its tile -- measured 0.40 -> 0.36 ms per 32 768 walkers against chunks of 32 with
// one CTA per SM); rows are padded by 4 doubles so both fragment loads are bank-conflict free.  Both operands have a leading dimension that is a multiple
// of 8 doubles and are zero padded beyond K (K1 writes the padding), so there is no K tail.
#include "jx_common.cuh"

namespace {

#ifndef K2_BK
#define K2_BK 16
#endif
#ifndef K2_CTAS
#define K2_CTAS 2
#endif
constexpr int BM = 128, BN = 64, BK = K2_BK, LDS = BK + 4, STAGES = 3;
constexpr int K2_THREADS = 256;
constexpr size_t K2_SMEM = (size_t)STAGES * (BM + BN) * LDS * sizeof(double);

JX_D void cp_async16(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
JX_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
JX_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

JX_D void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// A: [M, lda] row-major, B: [N, ldb] row-major (= K x N column-major), C: [M, ldc]
__global__ void __launch_bounds__(K2_THREADS, K2_CTAS)
k2_dgemm_nt_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                   double* __restrict__ C, int ldc, int M, int N, int Kpad) {
    extern __shared__ __align__(16) double k2_smem[];
    double* As = k2_smem;                                  // [STAGES][BM][LDS]
    double* Bs = k2_smem + (size_t)STAGES * BM * LDS;      // [STAGES][BN][LDS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;               // 4 x 2 warps
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;   // N tiles fastest: the CTAs sharing an A tile run together
    const int nchunks = Kpad / BK + ((Kpad % BK) ? 1 : 0);

    auto load_stage = [&](int stage, int chunk) {
        const int k0 = chunk * BK;
        // A tile: 128 rows x 32 doubles = 128 x 16 float4-sized pieces -> 2048 pieces / 256 threads
#pragma unroll
        for (int it = 0; it < (BM * BK / 2) / K2_THREADS; ++it) {
            int piece = it * K2_THREADS + tid;
            int row = piece / (BK / 2), col = (piece % (BK / 2)) * 2;
            bool ok = (m0 + row < M) && (k0 + col < Kpad);
            const double* src = A + (size_t)(ok ? m0 + row : 0) * lda + (ok ? k0 + col : 0);
            cp_async16(As + ((size_t)stage * BM + row) * LDS + col, src, ok);
        }
#pragma unroll
        for (int it = 0; it < (BN * BK / 2) / K2_THREADS; ++it) {
            int piece = it * K2_THREADS + tid;
            int row = piece / (BK / 2), col = (piece % (BK / 2)) * 2;
            bool ok = (n0 + row < N) && (k0 + col < Kpad);
            const double* src = B + (size_t)(ok ? n0 + row : 0) * ldb + (ok ? k0 + col : 0);
            cp_async16(Bs + ((size_t)stage * BN + row) * LDS + col, src, ok);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nchunks) load_stage(s, s);
        cp_async_commit();
    }
    const int frow = lane >> 2, fk = lane & 3;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        // prefetch the chunk that reuses the slot consumed in the previous iteration
        int next = chunk + STAGES - 1;
        if (next < nchunks) load_stage(next % STAGES, next);
        cp_async_commit();
        const double* as = As + ((size_t)(chunk % STAGES) * BM + wm * 32 + frow) * LDS + fk;
        const double* bs = Bs + ((size_t)(chunk % STAGES) * BN + wn * 32 + frow) * LDS + fk;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = as[(size_t)i * 8 * LDS + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = bs[(size_t)j * 8 * LDS + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: lane owns C[row = lane/4][col = 2*(lane%4) + {0,1}] of each 8x8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + wm * 32 + i * 8 + frow;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + wn * 32 + j * 8 + 2 * fk;
            if (n < N) C[(size_t)m * ldc + n] = acc[i][j][0];
            if (n + 1 < N) C[(size_t)m * ldc + n + 1] = acc[i][j][1];
        }
    }
}

}  // namespace

// per device, at jx_create
cudaError_t jx_gemm_configure() {
    return cudaFuncSetAttribute(k2_dgemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM);
}

// C[M, N] = A[M, Kpad] . B[N, Kpad]^T; lda/ldb even, A and B zero padded up to Kpad (a multiple of 8)
cudaError_t jx_launch_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N,
                              int Kpad, cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    k2_dgemm_nt_kernel<<<grid, K2_THREADS, K2_SMEM, st>>>(A, lda, B, ldb, C, ldc, M, N, Kpad);
    return cudaGetLastError();
}

// pp: [W, ldk] (ldk = round_up(nr, 8), zero padded), op: [nout, ldk] zero padded, out: [W, nout]
cudaError_t jx_launch_project(const jx_dev& d, const double* pp, int W, const double* op, int nout,
                              double* out, cudaStream_t st) {
    return jx_launch_gemm_nt(pp, d.nrp, op, d.nrp, out, nout, W, nout, d.nrp, st);
}
