// K3 -- Compton-y map synthesis and beam convolution, one persistent CTA per SM looping over walkers; the
// whole map pipeline of a walker lives in shared memory.
//
// Replaces, per walker, reference joxsz_funcs.py:462 (y_2d = f(d_mat)) and :464 (fftconvolve 'same').
//
// Geometry facts used (checked at pack time, joxsz_b200/operators.py): the map side N is odd and the
// Compton-y map, the beam and the filter are symmetric under x -> -x, y -> -y about the centre pixel.
// Re-centred on the origin every 2-D DFT of the stage is therefore a REAL cosine transform of the
// quarter plane (H = N/2+1 rows/cols), and two real even sequences ride in one complex FFT (real part /
// imaginary part) with no post-processing.  Per walker:
//
//   A  rows    Z[u,:]  (spline pieces evaluated on the fly)  --FFT256-->  xs[u, kx]      kx = 0..P/2
//   B  columns xs[:, kx]  (*) beam in (y offset, kx), 55 taps, directly           (small beams: the shipped case)
//              or  xs[:, kx] --FFT256--> * bhat[ky, kx] --FFT256--> xs[u, kx]     (any beam, cyclic length P)
//   C  rows    xs[u, :]  --FFT256--> conv_c[u, v]                = fftconvolve(y_2d, beam,'same')*step^2
// and the kernel writes conv_c on u <= v (the convolved map is symmetric under x <-> y as well), packed row-major,
// H (H + 1) / 2 doubles per walker, straight from the registers of phase C.  The remaining steps are batched over
// walkers outside this kernel: the exact length-N circular filter (N = 171 = 9*19 for the shipped cluster, no fast
// transform) restricted to the consumed row is one DMMA GEMM with a constant operator (k7_filter.cu), then the
// conversion / chi^2 tail (k5_tail.cu).
// The per-walker inputs (spline coefficients) arrive by TMA bulk copy (cp.async.bulk + mbarrier),
// double buffered against the previous walker's compute; the per-thread constants of the walker loop
// (synthesis-table entries, beam taps, FFT twiddles) sit in tensor memory (jx_tmem.cuh).
#include "k3_common.cuh"
#include "jx_tmem.cuh"
#include <stdlib.h>

// Developer instrumentation (scripts/k3_phase_clocks.py builds a private copy of the library with
// -DJX_K3_CLOCKS): SM cycles per phase, summed over CTAs and walkers.  Never compiled into the shipped library.
#ifdef JX_K3_CLOCKS
__device__ unsigned long long jx_k3_clk[8];
#define K3_CLK_DECL long long k3_t0 = clock64()
#define K3_CLK(i) do { if (threadIdx.x == 0) { long long k3_t1 = clock64(); atomicAdd(&jx_k3_clk[i], (unsigned long long)(k3_t1 - k3_t0)); k3_t0 = k3_t1; } } while (0)
extern "C" int jx_debug_k3_clocks(unsigned long long* out8) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, jx_k3_clk, sizeof(jx_k3_clk));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(jx_k3_clk, z, sizeof(z));
    return e == cudaSuccess ? 0 : -1;
}
#else
#define K3_CLK_DECL
#define K3_CLK(i)
#endif

namespace {

// CTA sizes, largest first; the first whose exchange buffers fit next to the map in shared memory is used.
// 512 threads = 16 warps x 3 nine-thread FFT groups (43 row pairs -> 1 round, 65 column pairs -> 2 rounds) at
// 128 registers per thread.
constexpr int K3_NT_A = 512, K3_NT_B = 384, K3_NT_C = 256;
constexpr int K3_P = 256;
constexpr int K3_Q = K3_P / 2 + 1;       // 129
constexpr int K3_XS = 130;               // row pitch of xs (doubles): XS/2 odd -> conflict-free column walks
// direct phase B: taps per side incl. the centre, rows per thread (4 warp groups), table pitch
constexpr int K3_NB = 28, K3_UB = 22, K3_BMP = JX_BMIX_PITCH;

struct k3_smem_layout {
    size_t tw, xbuf, xs, coef, nyqt, mbar, tmem, total;
};

__host__ __device__ inline k3_smem_layout k3_layout(const jx_dev& d, int hp8, int nthreads) {
    k3_smem_layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
    L.tw = take((size_t)JX_XE_ROWS * 32 * sizeof(double2));
    L.xbuf = take((size_t)(nthreads / 32) * 3 * JX_XE_ELEMS * sizeof(double2));
    L.xs = take((size_t)hp8 * K3_XS * sizeof(double));
    L.coef = take((size_t)2 * d.ncoef * sizeof(double));
    L.nyqt = take(K3_NB * sizeof(double));             // taps of the Nyquist column (direct phase B)
    L.mbar = take(2 * sizeof(uint64_t));
    L.tmem = take(sizeof(uint32_t));
    L.total = o;
    return L;
}

template <int NT, bool BDIRECT>
__global__ void __launch_bounds__(NT, 1) k3_szmap_kernel(const __grid_constant__ k3_args a) {
    extern __shared__ __align__(128) unsigned char k3_raw[];
    const jx_dev& d = a.d;
    const int H = d.nh, hp8 = d.hp8;
    const k3_smem_layout L = k3_layout(d, hp8, NT);
    double2* tw_s = reinterpret_cast<double2*>(k3_raw + L.tw);
    double2* xbuf_all = reinterpret_cast<double2*>(k3_raw + L.xbuf);
    double* xs = reinterpret_cast<double*>(k3_raw + L.xs);
    double* coef_s = reinterpret_cast<double*>(k3_raw + L.coef);
    [[maybe_unused]] double* nyqt = reinterpret_cast<double*>(k3_raw + L.nyqt);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(k3_raw + L.mbar);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // FFT groups: every sequence the kernel transforms is even, so nine threads carry one length-256 transform
    // (jx_fft.cuh) and a warp holds three groups; lanes 27..31 idle along (no shared-memory access at all: a fifth
    // lane group in a quarter-warp would collide with the banks of group 2)
    constexpr int NW = NT / 32;
    const bool lane_on = lane < 27;
    const int fg = lane_on ? lane / 9 : 2;
    const int t = lane_on ? lane - 9 * fg : lane - 27;
    const bool t_edge = t == 0 || t == 8;                   // threads whose outputs beyond index 128 repeat earlier ones
    double2* xbuf = xbuf_all + (size_t)(warp * 3 + fg) * JX_XE_ELEMS;
    const double2* tw_l = tw_s + lane;
    const uint32_t coef_bytes = (uint32_t)(d.ncoef * sizeof(double));

    // ---- one-time set-up of the CTA
    if constexpr (!BDIRECT) {                                // (the direct variant keeps its twiddles in tensor memory)
        for (int i = tid; i < JX_XE_ROWS * 32; i += NT) {    // per-lane twiddles [k2][lane]: w256^(t(lane) k2)
            const int l = i & 31, tl = l < 27 ? l % 9 : l - 27;
            fft256_make_twiddle((i >> 5) * 16 + tl, tw_s[i]);
        }
    }
    for (int i = tid; i < hp8 * K3_XS; i += NT) xs[i] = 0.0;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // Direct variant: the per-thread constants of the walker loop -- this thread's 8 entries of the synthesis table
    // (columns 0..31 of its slot), its 28 beam taps (32..87) and its 8 FFT twiddles (96..127) -- live in tensor
    // memory (unused otherwise: no tcgen05.mma in an FP64 kernel), written once here and read back with tcgen05.ld
    // every walker instead of 36 loads from L2 and 32 from shared memory.
    // Thread i of warp w owns TMEM lane 32 (w % 4) + i; the four warps of a lane quarter take 128 columns each.
    [[maybe_unused]] uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k3_raw + L.tmem);
    [[maybe_unused]] uint32_t tm_base = 0, tm_mine = 0;
    if constexpr (BDIRECT) {
        if (warp == 0) tmem_alloc(tmem_slot, 512);
        tmem_fence_before_sync();
    }
    __syncthreads();
    if constexpr (BDIRECT) {
        tmem_fence_after_sync();
        tm_base = *tmem_slot;
        tm_mine = tm_base + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (uint32_t)(warp >> 2);
        uint32_t r[64];
        const int4* tab = reinterpret_cast<const int4*>(d.synth);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = k * NT + tid;
            const int4 e = i < d.nsynth ? __ldg(tab + i) : make_int4(0, 0, 0xffff0000, 0);
            r[4 * k] = (uint32_t)e.x; r[4 * k + 1] = (uint32_t)e.y; r[4 * k + 2] = (uint32_t)e.z; r[4 * k + 3] = (uint32_t)e.w;
        }
        tmem_st32(tm_mine, reinterpret_cast<uint32_t(&)[32]>(r));
        tmem_wait_st();
        const int kxc = 32 * (warp & 3) + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) {           // taps 0..27 of column kxc (+ 4 pads)
            const double v = j < K3_NB ? __ldg(d.bmix + j * K3_BMP + kxc) : 0.0;
            r[2 * j] = (uint32_t)__double2loint(v); r[2 * j + 1] = (uint32_t)__double2hiint(v);
        }
        tmem_st64(tm_mine + 32, r);
        tmem_wait_st();
#pragma unroll
        for (int k2 = 1; k2 < JX_XE_ROWS; ++k2) {            // w256^(t k2), k2 = 1..8
            double2 tw;
            fft256_make_twiddle(k2 * 16 + t, tw);
            r[4 * (k2 - 1)] = (uint32_t)__double2loint(tw.x); r[4 * (k2 - 1) + 1] = (uint32_t)__double2hiint(tw.x);
            r[4 * (k2 - 1) + 2] = (uint32_t)__double2loint(tw.y); r[4 * (k2 - 1) + 3] = (uint32_t)__double2hiint(tw.y);
        }
        tmem_st32(tm_mine + 96, reinterpret_cast<uint32_t(&)[32]>(r));
        tmem_wait_st();
        if (tid < K3_NB) nyqt[tid] = __ldg(d.bmix + tid * K3_BMP + 128);
        __syncthreads();
    }

    const int w_first = blockIdx.x;
    if (tid == 0 && w_first < a.W) {
        mbar_expect_tx(&mbar[0], coef_bytes);
        tma_bulk_g2s(coef_s, a.coef + (size_t)w_first * d.ncoef, coef_bytes, &mbar[0]);
    }

    int it = 0;
    // status bits of the walker one iteration ahead, so that the load is never waited for.  Only the bits the profile
    // kernel wrote (final before this kernel starts) decide the skip: the X-ray kernel may still be OR-ing
    // JX_FLAG_XNONPOS into the word on the side stream, and a bit that changes under the readers would make the
    // decision differ between the warps of a CTA (mismatched barriers).  The tail kernel sees every bit.
    constexpr uint32_t K3_SKIP_BITS = ~(uint32_t)JX_FLAG_XNONPOS;
    uint32_t flag_next = (a.flags && w_first < a.W) ? (a.flags[w_first] & K3_SKIP_BITS) : 0u;
    for (int w = w_first; w < a.W; w += gridDim.x, ++it) {
        const int buf = it & 1;
        const double* cf = coef_s + (size_t)buf * d.ncoef;
        // prefetch the next walker's coefficients into the other buffer (its last readers finished
        // phase A of the previous iteration, several block barriers ago)
        {
            const int wn = w + gridDim.x;
            if (tid == 0 && wn < a.W) {
                mbar_expect_tx(&mbar[buf ^ 1], coef_bytes);
                tma_bulk_g2s(coef_s + (size_t)(buf ^ 1) * d.ncoef, a.coef + (size_t)wn * d.ncoef, coef_bytes,
                             &mbar[buf ^ 1]);
            }
        }
        const bool skip = flag_next != 0u;
        {
            const int wn = w + gridDim.x;
            flag_next = (a.flags && wn < a.W) ? (a.flags[wn] & K3_SKIP_BITS) : 0u;
        }
        mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));
        if (skip) {                         // block-uniform: the tail kernel writes -inf for flagged walkers
            __syncthreads();                // nobody still polls this mbarrier when thread 0 re-arms it
            continue;
        }

        double re[16], im[16];
        K3_CLK_DECL;

        // ================= phase A0: synthesise the quarter-plane map into xs[u, v].  The map is radial, so
        // pixel (u, v) also fills (v, u); the table lists u <= v in thread order (one coalesced 16-byte load
        // per pixel, all of a thread's loads in flight together).
        {
            constexpr int A0_UNROLL = 8;
            const int4* tab = reinterpret_cast<const int4*>(d.synth);
            for (int base = 0; base < d.nsynth; base += A0_UNROLL * NT) {
                int4 e[A0_UNROLL];
                if constexpr (BDIRECT) {         // nsynth <= 8 NT: one pass, entries from tensor memory
                    static_assert(A0_UNROLL == 8, "TMEM holds 8 table entries per thread");
                    uint32_t r[32];
                    tmem_ld32(tm_mine, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < A0_UNROLL; ++k) e[k] = make_int4((int)r[4 * k], (int)r[4 * k + 1], (int)r[4 * k + 2], (int)r[4 * k + 3]);
                } else {
#pragma unroll
                    for (int k = 0; k < A0_UNROLL; ++k) {
                        const int i = base + k * NT + tid;
                        e[k] = i < d.nsynth ? __ldg(tab + i) : make_int4(0, 0, 0xffff0000, 0);
                    }
                }
#pragma unroll
                for (int k = 0; k < A0_UNROLL; ++k) {
                    const int sg = e[k].z & 0xffff, u = (e[k].z >> 16) & 0xffff, v = e[k].w & 0xffff;
                    if (u != 0xffff) {
                        const double z = spline_eval(cf, d.nseg, sg, __hiloint2double(e[k].y, e[k].x));
                        xs[u * K3_XS + v] = z;
                        xs[v * K3_XS + u] = z;
                    }
                }
            }
        }
        __syncthreads();
        K3_CLK(0);

        // pass 1 of a row transform; the direct variant takes its twiddles from tensor memory (issued by tw_fetch
        // before the row loads, so that the two latencies overlap) instead of eight 16-byte shared-memory loads
        [[maybe_unused]] uint32_t twr[BDIRECT ? 32 : 1];
        auto tw_fetch = [&]() {
            if constexpr (BDIRECT) tmem_ld32(tm_mine + 96, reinterpret_cast<uint32_t(&)[32]>(twr));
        };
        auto row_pass1 = [&](double (&xr)[16], double (&xi)[16]) {
            if constexpr (BDIRECT) {
                tmem_wait_ld();
                double2 w[JX_XE_ROWS];
                w[0] = make_double2(1.0, 0.0);
#pragma unroll
                for (int k2 = 1; k2 < JX_XE_ROWS; ++k2)
                    w[k2] = make_double2(__hiloint2double((int)twr[4 * k2 - 3], (int)twr[4 * k2 - 4]),
                                         __hiloint2double((int)twr[4 * k2 - 1], (int)twr[4 * k2 - 2]));
                fft256e_pass1_w(t, xr, xi, w, xbuf, lane_on);
            } else {
                fft256e_pass1<32>(t, xr, xi, tw_l, xbuf, lane_on);
            }
        };

        // ================= phase A1: transform the rows along x, in place
        const int npair = (H + 1) >> 1;
        for (int base = warp * 3; base < npair; base += 3 * NW) {
            tw_fetch();
            const bool ok = lane_on && base + fg < npair;
            const int u0 = 2 * (base + fg < npair ? base + fg : npair - 1), u1 = u0 + 1;
            const bool has1 = u1 < H;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int f = fold256(t + 16 * j);
                double vr = 0.0, vi = 0.0;
                if (lane_on && f < H) {
                    vr = xs[u0 * K3_XS + f];
                    if (has1) vi = xs[u1 * K3_XS + f];
                }
                re[j] = vr; im[j] = vi;
            }
            row_pass1(re, im);
            __syncwarp();
            fft256_pass2(t, re, im, xbuf, lane_on);
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int n = t + 16 * rev16(p);                     // spectrum index held at position p
                if (ok && !(n > 128 && t_edge)) {
                    const int k = fold256(n);
                    xs[u0 * K3_XS + k] = re[p];
                    if (has1) xs[u1 * K3_XS + k] = im[p];
                }
            }
        }
        [[maybe_unused]] double tap[BDIRECT ? K3_NB : 1];
        [[maybe_unused]] const int bkx = 32 * (warp & 3) + lane;
        [[maybe_unused]] double ntap[BDIRECT ? 7 : 1];
        [[maybe_unused]] const int nyq_c = lane >> 3, nyq_u = warp * 6 + (lane & 7);
        [[maybe_unused]] const bool nyq_on = (lane & 7) < 6 && nyq_u < H;
        static_assert(K3_NB == 28, "the Nyquist column splits 28 taps into 4 chunks of 7");
        __syncthreads();
        K3_CLK(1);

        // ================= phase B: beam convolution along y, in place, for every x frequency kx
        if constexpr (BDIRECT) {
            // Small beam (at most K3_NB taps each side) and H <= 4 K3_UB: direct convolution in the mixed domain,
            //   xs[u, kx] <- sum_j bmix[|j|, kx] ext(xs)[u - j, kx],
            // about the flop count of the two column FFTs but plain FMA streams on 22 independent accumulators: no
            // exchange, every lane busy, and the 129 columns split evenly (the FFT form needs 65 column pairs on
            // 48 group slots, i.e. a second, mostly idle round).  Lane = column (conflict-free row segments), the
            // four warp groups take a quarter of the rows each; the Nyquist column kx = 128 is shared out by rows.
            static_assert(NT == 512, "direct phase B is laid out for 16 warps");
            const int kx = bkx, u0 = (warp >> 2) * K3_UB;
            {   // this thread's taps, from tensor memory; the Nyquist taps from a 28-entry shared table
                uint32_t r[64];
                tmem_ld64(tm_mine + 32, r);
#pragma unroll
                for (int i = 0; i < 7; ++i) ntap[i] = nyqt[7 * nyq_c + i];
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < K3_NB; ++j) tap[j] = __hiloint2double((int)r[2 * j + 1], (int)r[2 * j]);
            }
            // Nyquist column kx = 128 first: warp w owns rows 6 w .. 6 w + 5, lane = (row slot, chunk of 7 taps)
            double nyq = 0.0;
            if (nyq_on) {
                const double* col = xs + 128;
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const int j = 7 * nyq_c + i, ua = nyq_u - j < 0 ? j - nyq_u : nyq_u - j, ub = nyq_u + j;
                    const double xa = col[ua * K3_XS], xb = j > 0 && ub < H ? col[ub * K3_XS] : 0.0;   // ua < H always
                    nyq = fma(ntap[i], xa + xb, nyq);
                }
            }
            nyq += __shfl_xor_sync(0xffffffffu, nyq, 8);
            nyq += __shfl_xor_sync(0xffffffffu, nyq, 16);
            double acc[K3_UB];
#pragma unroll
            for (int k = 0; k < K3_UB; ++k) acc[k] = 0.0;
#pragma unroll
            for (int ii = 0; ii < K3_UB + 2 * (K3_NB - 1); ++ii) {
                const int up = u0 - (K3_NB - 1) + ii, ua = up < 0 ? -up : up;
                const double x = ua < H ? xs[ua * K3_XS + kx] : 0.0;
#pragma unroll
                for (int k = 0; k < K3_UB; ++k) {
                    const int j = ii - (K3_NB - 1) - k < 0 ? k + (K3_NB - 1) - ii : ii - (K3_NB - 1) - k;
                    if (j < K3_NB) acc[k] = fma(tap[j], x, acc[k]);
                }
            }
            __syncthreads();                      // every input has been read: the columns may be overwritten
#pragma unroll
            for (int k = 0; k < K3_UB; ++k)
                if (u0 + k < H) xs[(u0 + k) * K3_XS + kx] = acc[k];
            if (nyq_on && nyq_c == 0) xs[nyq_u * K3_XS + 128] = nyq;
        } else {
        // columns: cyclic convolution through two FFTs per column pair (any beam / map size the kernel admits)
        const int ncpair = (K3_Q + 1) >> 1;
        for (int base = warp * 3; base < ncpair; base += 3 * NW) {
            const bool ok = lane_on && base + fg < ncpair;
            const int cp = base + fg < ncpair ? base + fg : ncpair - 1;
            const int kx = 2 * cp;
            const bool has1 = kx + 1 < K3_Q;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int f = fold256(t + 16 * j);
                double2 v = make_double2(0.0, 0.0);
                if (lane_on && f < H) v = *reinterpret_cast<const double2*>(xs + f * K3_XS + kx);
                re[j] = v.x; im[j] = has1 ? v.y : 0.0;
            }
            fft256e_pass1<32>(t, re, im, tw_l, xbuf, lane_on);
            __syncwarp();
            fft256_pass2(t, re, im, xbuf, lane_on);
            __syncwarp();
            double re2[16], im2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {           // spectrum * beam, back to natural register order
                // beam spectrum of this column pair in thread order [cp][p][t]: 144 contiguous bytes per group
                const double2 bh = __ldg(d.bhat_sw + ((size_t)cp * 16 + rev16(j)) * 9 + t);
                re2[j] = re[rev16(j)] * bh.x;
                im2[j] = im[rev16(j)] * bh.y;
            }
            fft256e_pass1<32>(t, re2, im2, tw_l, xbuf, lane_on);
            __syncwarp();
            fft256_pass2(t, re2, im2, xbuf, lane_on);
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int n = t + 16 * rev16(p);
                const int u = fold256(n);
                if (ok && u < H && !(n > 128 && t_edge))
                    *reinterpret_cast<double2*>(xs + u * K3_XS + kx) = make_double2(re2[p], has1 ? im2[p] : 0.0);
            }
        }
        }
        __syncthreads();
        K3_CLK(2);

        // ================= phase C: rows back to pixel space; conv_c[u, v] for v >= u goes to the packed triangle
        // of this walker (row u starts at u H - u (u - 1) / 2), 72-byte runs per (row, register position)
        double* tri_w = a.tri + (size_t)w * d.ktri;
        const bool tapq = a.convq != nullptr;                  // parity tap: also keep the full quarter plane
        for (int base = warp * 3; base < npair; base += 3 * NW) {
            tw_fetch();
            const bool ok = lane_on && base + fg < npair;
            const int u0 = 2 * (base + fg < npair ? base + fg : npair - 1), u1 = u0 + 1;
            const bool has1 = u1 < H;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int f = fold256(t + 16 * j);
                re[j] = lane_on ? xs[u0 * K3_XS + f] : 0.0;
                im[j] = lane_on && has1 ? xs[u1 * K3_XS + f] : 0.0;
            }
            row_pass1(re, im);
            __syncwarp();
            fft256_pass2(t, re, im, xbuf, lane_on);
            double* tri0 = tri_w + (u0 * H - ((u0 * (u0 - 1)) >> 1) - u0);     // + v
            double* tri1 = tri0 + (H - u0 - 1);                                 // row u1: off(u0) + (H - u0) - u1
            if (tapq) __syncwarp();   // every lane of the warp has read its row pair before the in-place stores
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int n = t + 16 * rev16(p);
                const int v = fold256(n);
                if (ok && v < H && !(n > 128 && t_edge)) {
                    if (v >= u0) tri0[v] = re[p];
                    if (has1 && v >= u1) tri1[v] = im[p];
                    if (tapq) {
                        xs[u0 * K3_XS + v] = re[p];
                        if (has1) xs[u1 * K3_XS + v] = im[p];
                    }
                }
            }
        }
        __syncthreads();        // all reads of xs are done: the next walker's synthesis (or the tap) may proceed
        K3_CLK(3);

        if (tapq) {
            double* cq = a.convq + (size_t)w * H * H;
            for (int i = tid; i < H * H; i += NT) cq[i] = xs[(i / H) * K3_XS + (i % H)];
            // the next iteration's first barrier is after its synthesis writes: order the tap reads before them
            __syncthreads();
        }
    }
    if constexpr (BDIRECT) {
        tmem_fence_before_sync();
        __syncthreads();
        if (warp == 0) tmem_dealloc(tm_base, 512);
    }
}

// ---- parity taps (not on the production path) -------------------------------------------------

// y_2d[w, y, x] = spline at d_mat[y, x]: quarter-plane pixel (|y-c|, |x-c|)
__global__ void k3_tap_y2d_kernel(jx_dev d, const double* coef, int W, double* y2d) {
    const int w = blockIdx.y;
    const int N = d.nmap, c = N / 2, H = d.nh;
    const double* cf = coef + (size_t)w * d.ncoef;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * N; i += gridDim.x * blockDim.x) {
        int y = i / N, x = i % N;
        int u = y < c ? c - y : y - c, v = x < c ? c - x : x - c;
        y2d[(size_t)w * N * N + i] = spline_eval(cf, d.nseg, d.seg16[u * H + v], d.dx[u * H + v]);
    }
}

__global__ void k3_tap_expand_kernel(jx_dev d, const double* convq, int W, double* full) {
    const int w = blockIdx.y;
    const int N = d.nmap, c = N / 2, H = d.nh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * N; i += gridDim.x * blockDim.x) {
        int y = i / N, x = i % N;
        int u = y < c ? c - y : y - c, v = x < c ? c - x : x - c;
        full[(size_t)w * N * N + i] = convq[(size_t)w * H * H + u * H + v];
    }
}

// map_out quarter plane by dense cosine transforms (tap only; O(H^3) per walker, one CTA per walker):
//   Chat = Cw conv Cw^T,  F = Chat * filt,  out = 1/N^2 Cw^T-style inverse on both axes
__global__ void k3_tap_mapout_kernel(jx_dev d, const double* costab, const double* convq, int W, double* mapout,
                                     double* scratch) {
    const int w = blockIdx.x;
    const int N = d.nmap, c = N / 2, H = d.nh;
    const double* cq = convq + (size_t)w * H * H;
    double* s1 = scratch + (size_t)w * 2 * H * H;
    double* s2 = s1 + (size_t)H * H;
    // s1[u, kx] = sum_v conv[u,v] w_v cos(kx v)
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
        int u = i / H, kx = i % H;
        double s = 0.0;
        for (int v = 0; v < H; ++v) s += cq[u * H + v] * (v ? 2.0 : 1.0) * costab[(kx * v) % N];
        s1[i] = s;
    }
    __syncthreads();
    // s2[ky, kx] = filt[ky,kx] * sum_u w_u cos(ky u) s1[u, kx]
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
        int ky = i / H, kx = i % H;
        double s = 0.0;
        for (int u = 0; u < H; ++u) s += s1[u * H + kx] * (u ? 2.0 : 1.0) * costab[(ky * u) % N];
        s2[i] = s * d.filt_q[i];
    }
    __syncthreads();
    // s1[ky, v] = sum_kx w_kx cos(kx v) s2[ky, kx]
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
        int ky = i / H, v = i % H;
        double s = 0.0;
        for (int kx = 0; kx < H; ++kx) s += s2[ky * H + kx] * (kx ? 2.0 : 1.0) * costab[(kx * v) % N];
        s1[i] = s;
    }
    __syncthreads();
    // s2[u, v] = 1/N^2 sum_ky w_ky cos(ky u) s1[ky, v]
    const double inv = 1.0 / ((double)N * (double)N);
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
        int u = i / H, v = i % H;
        double s = 0.0;
        for (int ky = 0; ky < H; ++ky) s += s1[ky * H + v] * (ky ? 2.0 : 1.0) * costab[(ky * u) % N];
        s2[i] = s * inv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
        int y = i / N, x = i % N;
        int u = y < c ? c - y : y - c, v = x < c ? c - x : x - c;
        mapout[(size_t)w * N * N + i] = s2[u * H + v];
    }
}

}  // namespace

static int k3_pick_threads(const jx_dev& d) {
    const size_t cap = 232448;               // 227 KB of shared memory per CTA on sm_100
    if (const char* e = getenv("JX_K3_THREADS")) {   // developer knob for A/B measurements
        const int nt = atoi(e);
        if ((nt == K3_NT_A || nt == K3_NT_B || nt == K3_NT_C) && k3_layout(d, d.hp8, nt).total <= cap) return nt;
    }
    if (k3_layout(d, d.hp8, K3_NT_A).total <= cap) return K3_NT_A;
    if (k3_layout(d, d.hp8, K3_NT_B).total <= cap) return K3_NT_B;
    return K3_NT_C;
}

// direct y convolution: beam and map small enough for the register tiling, 16 warps; JX_K3_BFFT=1 (read once, at
// jx_create) forces the FFT form
bool jx_szmap_direct_ok(const jx_dev& d) {
    if (const char* e = getenv("JX_K3_BFFT")) if (atoi(e)) return false;
    return k3_pick_threads(d) == K3_NT_A && d.bmix && d.npad == K3_P && d.nbeam <= K3_NB && d.nh <= 4 * K3_UB &&
           d.nsynth <= 8 * K3_NT_A;
}

template <class F>
static auto k3_dispatch(const jx_dev& d, F&& f) {
    const int nt = k3_pick_threads(d);
    if (nt == K3_NT_A) return d.k3_direct ? f(k3_szmap_kernel<K3_NT_A, true>, nt) : f(k3_szmap_kernel<K3_NT_A, false>, nt);
    if (nt == K3_NT_B) return f(k3_szmap_kernel<K3_NT_B, false>, nt);
    return f(k3_szmap_kernel<K3_NT_C, false>, nt);
}

cudaError_t jx_szmap_configure(const jx_dev& d) {
    return k3_dispatch(d, [&](auto kern, int nt) {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_layout(d, d.hp8, nt).total);
    });
}

size_t jx_szmap_smem_bytes(const jx_dev& d) { return k3_layout(d, d.hp8, k3_pick_threads(d)).total; }

cudaError_t jx_launch_szmap(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                            double* convq, double* tri, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k3_args a;
    a.d = d; a.coef = coef; a.flags = flags; a.W = W; a.convq = convq; a.tri = tri; a.scratch = nullptr; a.scratch2 = nullptr;
    const int grid = W < sm_count ? W : sm_count;
    return k3_dispatch(d, [&](auto kern, int nt) {
        kern<<<grid, nt, k3_layout(d, d.hp8, nt).total, st>>>(a);
        return cudaGetLastError();
    });
}

cudaError_t jx_launch_tap_y2d(const jx_dev& d, const double* coef, int W, double* y2d, cudaStream_t st) {
    for (int w0 = 0; w0 < W; w0 += 32768) {
        int wc = W - w0 < 32768 ? W - w0 : 32768;
        dim3 grid(32, wc);
        k3_tap_y2d_kernel<<<grid, 256, 0, st>>>(d, coef + (size_t)w0 * d.ncoef, wc,
                                                 y2d + (size_t)w0 * d.nmap * d.nmap);
    }
    return cudaGetLastError();
}

cudaError_t jx_launch_tap_expand(const jx_dev& d, const double* convq, int W, double* conv2d, cudaStream_t st) {
    for (int w0 = 0; w0 < W; w0 += 32768) {
        int wc = W - w0 < 32768 ? W - w0 : 32768;
        dim3 grid(32, wc);
        k3_tap_expand_kernel<<<grid, 256, 0, st>>>(d, convq + (size_t)w0 * d.nh * d.nh, wc,
                                                    conv2d + (size_t)w0 * d.nmap * d.nmap);
    }
    return cudaGetLastError();
}

// scratch: [W, 2, H, H] doubles owned by the caller
cudaError_t jx_launch_tap_mapout(const jx_dev& d, const double* convq, int W, double* mapout, double* scratch,
                                 cudaStream_t st) {
    k3_tap_mapout_kernel<<<W, 256, 0, st>>>(d, d.costab, convq, W, mapout, scratch);
    return cudaGetLastError();
}
