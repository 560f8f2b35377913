// K7 -- transfer-function filtering of the beam-convolved map as ONE dense FP64 contraction over all walkers.
//
// Replaces, for every walker of the batch, reference joxsz_funcs.py:466-467
//     map_out = real(ifft2(fft2(conv) * filtering))
// restricted to the part the likelihood consumes, map_out[N//2, N//2:] (:472).  The circular filter at the exact
// map size N (171 = 9 * 19 for the shipped cluster: no fast transform) is linear, and the convolved map is
// symmetric under x <-> y, x -> -x, y -> -y, so the H = N//2 + 1 consumed values are a fixed linear map of the
// H (H + 1) / 2 distinct pixels u <= v of the quarter plane:
//
//     row[w, x] = sum_{u <= v} tri[w, (u, v)] R[(u, v), x]
//     R[(u, v), x] = sum_kx (hf[u,kx] cmat[v,kx] + [u != v] hf[v,kx] cmat[u,kx]) dinv[kx, x]     (built at jx_create)
//
// with hf = filter folded with the k_y cosine sum, cmat = w_v cos(2 pi kx v / N), dinv = inverse cosine transform
// of the row (joxsz_b200/operators.py).  `tri` is what the map kernel (k3_szmap.cu) writes: the convolved map
// fftconvolve(y_2d, beam, 'same') * step^2 (:464) on u <= v, row-major packed, zero padded to a multiple of 32.
// Half the flops of transforming every row of the quarter plane (K = H(H+1)/2 instead of H^2), and a regular
// GEMM with a constant, L2-resident B operand instead of a per-walker one.
//
// FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA): the 1e-6 absolute bar on the log-likelihood needs double
// operands and accumulation, and tcgen05 has no f64 kind.
//
// Tiling: CTA = 128 walkers x all 8 NT outputs, 16 warps, warp tile 8 x 8 NT (NT DMMA tiles; four warps per
// scheduler keep the FP64 tensor pipe fed across the per-chunk barrier), K in chunks
// of 32 through a 3-stage cp.async ring (rows padded to 36 doubles: both fragment loads take the minimum two
// wavefronts).  K is additionally split into contiguous parts of JX_FILTER_CPP chunks (grid.x) so that walker tiles
// x parts fills the SMs evenly; part p writes its partial sums to C + p * M * ldc and the tail kernel
// (k5_tail.cu) adds the parts in order -- deterministic, no atomics, independent of the batch size.
#include "jx_common.cuh"

namespace {

// Two tilings.  Narrow quarter planes (at most 11 DMMA column tiles, the shipped 171-pixel maps): 8 warps x (16 x 8 NT),
// K chunks of 16, 104 KB of shared memory -> two CTAs per SM, one computing while the other fills its ring or stores
// (measured at the shipped geometry: 0.74 -> 0.68 ms per 32 768 walkers, 0.84 of the DMMA peak).  Wide ones (255 / 511
// pixels) need the whole SM for one CTA: 16 warps x (8 x 8 NT), K chunks of 32.
constexpr int K7_BM = 128, K7_STAGES = 3;
constexpr int K7_NT_TWO_CTAS = 11;
template <int NT> struct k7_cfg {
    static constexpr bool two = NT <= K7_NT_TWO_CTAS;
    static constexpr int WARPS = two ? 8 : 16, BK = two ? 16 : 32, CTAS = two ? 2 : 1;
    static constexpr int LDS = BK + 4, THREADS = 32 * WARPS, MT = K7_BM / (8 * WARPS);
    static constexpr size_t SMEM = (size_t)K7_STAGES * (K7_BM + 8 * NT) * LDS * sizeof(double);
};

JX_D void k7_cp16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
JX_D void k7_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
JX_D void k7_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

JX_D void k7_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NT>
constexpr size_t k7_smem_bytes() { return k7_cfg<NT>::SMEM; }

// A: [M, lda] packed convolved maps (lda = K rounded up to 32, zero padded), B: [ldc, lda] = R^T zero padded,
// C: [kparts][M][ldc]; blockIdx.z selects a block of 8 NT output columns (quarter planes wider than 136 pixels)
template <int NT>
__global__ void __launch_bounds__(k7_cfg<NT>::THREADS, k7_cfg<NT>::CTAS)
k7_filter_gemm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int lda, int ldc,
                      int M, int nchunks_total, int chunks_per_part) {
    extern __shared__ __align__(16) double k7_smem[];
    constexpr int BN = 8 * NT;
    constexpr int K7_BK = k7_cfg<NT>::BK, K7_LDS = k7_cfg<NT>::LDS, K7_THREADS = k7_cfg<NT>::THREADS, K7_MT = k7_cfg<NT>::MT;
    double* As = k7_smem;                                          // [STAGES][BM][LDS]
    double* Bs = k7_smem + (size_t)K7_STAGES * K7_BM * K7_LDS;     // [STAGES][BN][LDS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int part = blockIdx.x, m0 = blockIdx.y * K7_BM, n0 = blockIdx.z * BN;
    const int c_first = part * chunks_per_part;
    const int nchunks = min(chunks_per_part, nchunks_total - c_first);

    auto load_stage = [&](int stage, int chunk) {
        const int k0 = (c_first + chunk) * K7_BK;
#pragma unroll
        for (int it = 0; it < (K7_BM * K7_BK / 2) / K7_THREADS; ++it) {
            const int piece = it * K7_THREADS + tid;
            const int row = piece / (K7_BK / 2), col = (piece % (K7_BK / 2)) * 2;
            const bool ok = m0 + row < M;
            k7_cp16(As + ((size_t)stage * K7_BM + row) * K7_LDS + col, A + (size_t)(ok ? m0 + row : 0) * lda + k0 + col, ok);
        }
#pragma unroll
        for (int it = 0; it < (BN * K7_BK / 2 + K7_THREADS - 1) / K7_THREADS; ++it) {
            const int piece = it * K7_THREADS + tid;
            if (piece < BN * K7_BK / 2) {
                const int row = piece / (K7_BK / 2), col = (piece % (K7_BK / 2)) * 2;
                k7_cp16(Bs + ((size_t)stage * BN + row) * K7_LDS + col, B + (size_t)(n0 + row) * lda + k0 + col, true);
            }
        }
    };

    double acc[K7_MT][NT][2];
#pragma unroll
    for (int i = 0; i < K7_MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int s = 0; s < K7_STAGES - 1; ++s) {
        if (s < nchunks) load_stage(s, s);
        k7_commit();
    }
    const int frow = lane >> 2, fk = lane & 3;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        k7_wait<K7_STAGES - 2>();
        __syncthreads();
        const int next = chunk + K7_STAGES - 1;          // reuses the slot consumed in the previous iteration
        if (next < nchunks) load_stage(next % K7_STAGES, next);
        k7_commit();
        const double* as = As + ((size_t)(chunk % K7_STAGES) * K7_BM + warp * (8 * K7_MT) + frow) * K7_LDS + fk;
        const double* bs = Bs + ((size_t)(chunk % K7_STAGES) * BN + frow) * K7_LDS + fk;
#pragma unroll
        for (int kk = 0; kk < K7_BK; kk += 4) {
            double bf[NT], af[K7_MT];
#pragma unroll
            for (int i = 0; i < K7_MT; ++i) af[i] = as[(size_t)i * 8 * K7_LDS + kk];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = bs[(size_t)j * 8 * K7_LDS + kk];
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int i = 0; i < K7_MT; ++i) k7_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    k7_wait<0>();

    // lane owns C[row = lane / 4][col = 2 (lane % 4) + {0, 1}] of each 8 x 8 tile
    double* Cp = C + (size_t)part * M * ldc + n0;
#pragma unroll
    for (int i = 0; i < K7_MT; ++i) {
        const int m = m0 + warp * (8 * K7_MT) + i * 8 + frow;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j)
            *reinterpret_cast<double2*>(Cp + (size_t)m * ldc + j * 8 + 2 * fk) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

#define K7_FOR_NT(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17)

}  // namespace

// Output columns are handled in `nblk` blocks of 8 NT columns, NT <= 17 (register budget of the warp tile).
void jx_filter_tiling(int hp8, int* nblk, int* nt) {
    const int ntile = hp8 / 8;
    *nblk = (ntile + 16) / 17;
    *nt = (ntile + *nblk - 1) / *nblk;
}

// leading dimension of filt_op's rows-of-outputs and of the partial rows: 8 NT nblk >= hp8
int jx_filter_pitch(int hp8) {
    int nblk, nt;
    jx_filter_tiling(hp8, &nblk, &nt);
    return 8 * nt * nblk;
}

cudaError_t jx_filter_configure(const jx_dev& d) {
    const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
    int nblk, nt;
    jx_filter_tiling(d.hp8, &nblk, &nt);
    switch (nt) {
#define K7_CASE(n) case n: return cudaFuncSetAttribute(k7_filter_gemm_kernel<n>, at, (int)k7_smem_bytes<n>());
        K7_FOR_NT(K7_CASE)
#undef K7_CASE
        default: return cudaErrorInvalidValue;
    }
}

// Number of K parts: fixed by the geometry alone, never by the batch size, so that a walker's result does not depend
// on which batch it is evaluated in (the sampler's chains are bit-identical for any number of ranks): parts of
// JX_FILTER_CPP chunks of 32, at most 9 parts (longer parts for the larger maps).  117 chunks -> 9 parts of 13 for
// the shipped cluster: 256 walker tiles x 9 parts fill 148 SMs to 97 %.
int jx_filter_parts(const jx_dev& d) {
    const int nchunks = d.ktri / 32, p = (nchunks + JX_FILTER_CPP - 1) / JX_FILTER_CPP;      // parts are counted in 32-wide chunks
    return p < 9 ? p : 9;
}

// rowp[kparts][W][hpf] = partial sums of tri[W, ktri] . filt_op[hpf, ktri]^T
cudaError_t jx_launch_filter(const jx_dev& d, const double* tri, int W, double* rowp, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    // part boundaries in units of 32 packed pixels whatever the kernel's own chunk width: the summation order of a
    // walker's row (and with it every bit of its log-likelihood) does not depend on the tiling
    const int nchunks32 = d.ktri / 32, parts = jx_filter_parts(d), cpp32 = (nchunks32 + parts - 1) / parts;
    int nblk, nt;
    jx_filter_tiling(d.hp8, &nblk, &nt);
    dim3 grid(parts, (W + K7_BM - 1) / K7_BM, nblk);
    switch (nt) {
#define K7_CASE(n) case n: k7_filter_gemm_kernel<n><<<grid, k7_cfg<n>::THREADS, k7_smem_bytes<n>(), st>>>( \
                               tri, d.filt_op, rowp, d.ktri, d.hpf, W, d.ktri / k7_cfg<n>::BK, cpp32 * (32 / k7_cfg<n>::BK)); break;
        K7_FOR_NT(K7_CASE)
#undef K7_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
