// K3W -- the map stage of k3_szmap.cu (Compton-y map synthesis + beam convolution, cyclic length 256) with TWO
// walkers in flight per SM: the CTA's 16 warps are split into a transform group and a convolution group that work
// on different walkers at the same time.
//
// Why: in k3_szmap.cu all 16 warps walk through the phases together.  The row transforms (A1, C) are bound by the
// shared-memory exchange and leave the FP64 pipe idle half of the time; the y convolution (B) streams FMAs and leaves
// the shared-memory pipe idle (ncu: FP64 pipe 52 % busy, shared pipe 47 %).  Here
//
//   warps 0-7  (F group)   A0 synthesis + A1 row transforms of walker s+1, then C rows-back + store of walker s-1
//   warps 8-15 (M group)   B, the 55-tap y convolution of walker s
//
// run concurrently on two shared-memory maps, so the FMA stream of one walker fills the FP64 pipe while the other
// walker's transforms exchange.  What makes two maps fit (2 x 103 KB + coefficients = 213 KB of the 227 KB):
//   * no separate exchange buffers: a nine-thread transform exchanges inside the two map rows it has just loaded into
//     registers (row pitch 153 doubles: a row pair is exactly the 9 x 17 complex exchange tile, and consecutive pairs
//     are 16 bytes apart modulo the 128-byte bank line, which keeps the three groups of a warp conflict free);
//   * the y convolution is done by 256 threads in two passes of 22 rows per thread; the first pass's 22 results wait
//     in TENSOR MEMORY (tcgen05.st / tcgen05.ld, 44 columns per thread) until every input row has been read, then both
//     passes are written back in place.
// Tensor memory also holds every per-thread constant (F threads: 16 synthesis-table entries + 8 twiddles; M threads:
// 28 beam taps), as in k3_szmap.cu.  Hand-offs between the groups are named barriers (bar.arrive / bar.sync):
// full[b] = "map b holds the row spectra of a walker", bdone[b] = "map b holds its convolved spectra".
//
// Results are bit-identical to k3_szmap_kernel<512, true> (same arithmetic in the same order per output).
#include "k3_common.cuh"
#include "jx_tmem.cuh"
#include <stdlib.h>

#ifdef JX_K3_CLOCKS
__device__ unsigned long long jx_k3w_clk[8];
extern "C" int jx_debug_k3w_clocks(unsigned long long* out8) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, jx_k3w_clk, sizeof(jx_k3w_clk));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(jx_k3w_clk, z, sizeof(z));
    return e == cudaSuccess ? 0 : -1;
}
#define K3W_CLK_DECL long long k3_t0 = clock64()
#define K3W_CLK(i) do { if ((threadIdx.x & 255) == 0) { long long k3_t1 = clock64(); atomicAdd(&jx_k3w_clk[i], (unsigned long long)(k3_t1 - k3_t0)); k3_t0 = k3_t1; } } while (0)
#else
#define K3W_CLK_DECL
#define K3W_CLK(i)
#endif

// Developer experiment (scripts/k3_phase_clocks.py with JX_CLK_DEFS; -DJX_K3W_NO_STG drops the triangle stores of C):
// -DJX_K3W_NO_F / -DJX_K3W_NO_M build the kernel
// with one warp group doing no work of its own (it keeps the barrier protocol), to time the other group alone.  The
// results of such a build are meaningless.
#ifdef JX_K3W_NO_F
#define KW_F_FIRST(x) (1 << 20)
#else
#define KW_F_FIRST(x) (x)
#endif
#ifdef JX_K3W_NO_M
#define KW_M_ON (a.W < 0)
#else
#define KW_M_ON true
#endif

namespace {

constexpr int KW_NT = 512, KW_NG = 256;      // CTA size, threads per group
constexpr int KW_P = 256;
constexpr int KW_XS = 153;                   // row pitch (doubles): two rows = 9 x 17 complex = one exchange tile
constexpr int KW_NB = 28, KW_UB = 22;        // taps per side incl. the centre, rows per thread and pass
constexpr int KW_SYN = 16;                   // synthesis-table entries per F thread (64 TMEM columns)
constexpr int KW_BMP = JX_BMIX_PITCH;
// named barriers (0 is __syncthreads)
constexpr int BAR_F = 1, BAR_M = 2, BAR_FULL = 3, BAR_BDONE = 5;

struct kw_layout {
    size_t map, coef, nyqt, mbar, tmem, total;
    int rows;
};

__host__ __device__ inline kw_layout kw_make_layout(const jx_dev& d) {
    kw_layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
    L.rows = 2 * ((d.nh + 1) >> 1);
    L.map = take((size_t)2 * L.rows * KW_XS * sizeof(double));
    L.coef = take((size_t)2 * d.ncoef * sizeof(double));
    L.nyqt = take(KW_NB * sizeof(double));
    L.mbar = take(2 * sizeof(uint64_t));
    L.tmem = take(sizeof(uint32_t));
    L.total = o;
    return L;
}

JX_D void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(n) : "memory"); }
JX_D void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(KW_NT, 1) k3w_szmap_kernel(const __grid_constant__ k3_args a) {
    extern __shared__ __align__(128) unsigned char kw_raw[];
    const jx_dev& d = a.d;
    const int H = d.nh;
    const kw_layout L = kw_make_layout(d);
    double* maps = reinterpret_cast<double*>(kw_raw + L.map);
    const int map_stride = L.rows * KW_XS;
    double* coef_s = reinterpret_cast<double*>(kw_raw + L.coef);
    double* nyqt = reinterpret_cast<double*>(kw_raw + L.nyqt);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(kw_raw + L.mbar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kw_raw + L.tmem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_f = warp < 8;
    const int gw = warp & 7, gtid = tid & 255;            // warp / thread index inside the group
    const uint32_t coef_bytes = (uint32_t)(d.ncoef * sizeof(double));

    // ---- one-time set-up of the CTA
    for (int i = tid; i < 2 * map_stride; i += KW_NT) maps[i] = 0.0;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tm_base = *tmem_slot;
    // thread i of warp w owns TMEM lane 32 (w % 4) + i; the four warps of a lane quarter take 128 columns each
    const uint32_t tm_mine = tm_base + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (uint32_t)(warp >> 2);

    // nine-thread transform groups of the F warps (three per warp, lanes 27..31 idle along)
    const bool lane_on = lane < 27;
    const int fg = lane_on ? lane / 9 : 2;
    const int t = lane_on ? lane - 9 * fg : lane - 27;
    const bool t_edge = t == 0 || t == 8;

    if (is_f) {
        uint32_t r[64];
        const int4* tab = reinterpret_cast<const int4*>(d.synth);
#pragma unroll
        for (int k = 0; k < KW_SYN; ++k) {           // columns 0..63: this thread's synthesis-table entries
            const int i = k * KW_NG + gtid;
            const int4 e = i < d.nsynth ? __ldg(tab + i) : make_int4(0, 0, 0xffff0000, 0);
            r[4 * k] = (uint32_t)e.x; r[4 * k + 1] = (uint32_t)e.y; r[4 * k + 2] = (uint32_t)e.z; r[4 * k + 3] = (uint32_t)e.w;
        }
        tmem_st64(tm_mine, r);
        tmem_wait_st();
#pragma unroll
        for (int k2 = 1; k2 < JX_XE_ROWS; ++k2) {    // columns 64..95: w256^(t k2), k2 = 1..8
            double2 tw;
            fft256_make_twiddle(k2 * 16 + t, tw);
            r[4 * (k2 - 1)] = (uint32_t)__double2loint(tw.x); r[4 * (k2 - 1) + 1] = (uint32_t)__double2hiint(tw.x);
            r[4 * (k2 - 1) + 2] = (uint32_t)__double2loint(tw.y); r[4 * (k2 - 1) + 3] = (uint32_t)__double2hiint(tw.y);
        }
        tmem_st32(tm_mine + 64, reinterpret_cast<uint32_t(&)[32]>(r));
        tmem_wait_st();
    } else {
        uint32_t r[64];
        const int kxc = 32 * (gw & 3) + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) {               // columns 0..63: taps 0..27 of column kxc (+ 4 pads)
            const double v = j < KW_NB ? __ldg(d.bmix + j * KW_BMP + kxc) : 0.0;
            r[2 * j] = (uint32_t)__double2loint(v); r[2 * j + 1] = (uint32_t)__double2hiint(v);
        }
        tmem_st64(tm_mine, r);
        tmem_wait_st();
        if (gtid < KW_NB) nyqt[gtid] = __ldg(d.bmix + gtid * KW_BMP + 128);
    }
    const int w_first = blockIdx.x;
    if (tid == 0 && w_first < a.W) {
        mbar_expect_tx(&mbar[0], coef_bytes);
        tma_bulk_g2s(coef_s, a.coef + (size_t)w_first * d.ncoef, coef_bytes, &mbar[0]);
    }
    __syncthreads();

    // only the bits the profile kernel wrote decide the skip (see k3_szmap.cu): both groups take the same decisions
    constexpr uint32_t SKIP_BITS = ~(uint32_t)JX_FLAG_XNONPOS;
    const int npair = (H + 1) >> 1;
    const bool tapq = a.convq != nullptr;

    if (is_f) {
        // =====================================================================================  F group
        uint32_t twr[32];
        auto row_pass1 = [&](double (&xr)[16], double (&xi)[16], double2* xbuf, bool on) {
            tmem_wait_ld();
            double2 w[JX_XE_ROWS];
            w[0] = make_double2(1.0, 0.0);
#pragma unroll
            for (int k2 = 1; k2 < JX_XE_ROWS; ++k2)
                w[k2] = make_double2(__hiloint2double((int)twr[4 * k2 - 3], (int)twr[4 * k2 - 4]),
                                     __hiloint2double((int)twr[4 * k2 - 1], (int)twr[4 * k2 - 2]));
            fft256e_pass1_w(t, xr, xi, w, xbuf, on);
        };
        // C of the walker `pw` whose convolved spectra are in map `pb`
        auto rows_back = [&](int pw, int pb) {
            K3W_CLK_DECL;
            bar_sync(BAR_BDONE + pb, KW_NT);
            K3W_CLK(4);
            double* xs = maps + (size_t)pb * map_stride;
            double* tri_w = a.tri + (size_t)pw * d.ktri;
            double re[16], im[16];
            for (int base = KW_F_FIRST(gw * 3); base < npair; base += 24) {
                tmem_ld32(tm_mine + 64, twr);
                const bool on = lane_on && base + fg < npair;
                const int u0 = 2 * (base + fg < npair ? base + fg : npair - 1), u1 = u0 + 1;
                const bool has1 = u1 < H;
                double2* xbuf = reinterpret_cast<double2*>(xs + u0 * KW_XS);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int f = fold256(t + 16 * j);
                    re[j] = on ? xs[u0 * KW_XS + f] : 0.0;
                    im[j] = on && has1 ? xs[u1 * KW_XS + f] : 0.0;
                }
                __syncwarp();                    // the rows are in registers: their memory becomes the exchange tile
                row_pass1(re, im, xbuf, on);
                __syncwarp();
                fft256_pass2(t, re, im, xbuf, on);
                double* tri0 = tri_w + (u0 * H - ((u0 * (u0 - 1)) >> 1) - u0);     // + v
                double* tri1 = tri0 + (H - u0 - 1);
                if (tapq) __syncwarp();          // exchange tile read by every lane before the in-place tap stores
                // straight from the registers: staging the rows in shared memory for coalesced 256-byte stores was
                // measured slower (C phase 13.3 k -> 15.5 k cycles per walker: the extra shared-memory traffic costs more
                // than the partial-sector stores)
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const int n = t + 16 * rev16(p);
                    const int v = fold256(n);
                    if (on && v < H && !(n > 128 && t_edge)) {
#ifndef JX_K3W_NO_STG
                        if (v >= u0) tri0[v] = re[p];
                        if (has1 && v >= u1) tri1[v] = im[p];
#else
                        if (a.W < 0) { tri0[v] = re[p]; tri1[v] = im[p]; }     // experiment: C without its global stores
#endif
                        if (tapq) {
                            xs[u0 * KW_XS + v] = re[p];
                            if (has1) xs[u1 * KW_XS + v] = im[p];
                        }
                    }
                }
            }
            bar_sync(BAR_F, KW_NG);              // the map may be synthesised into again (or tapped)
            if (tapq) {
                double* cq = a.convq + (size_t)pw * H * H;
                for (int i = gtid; i < H * H; i += KW_NG) cq[i] = xs[(i / H) * KW_XS + (i % H)];
                bar_sync(BAR_F, KW_NG);
            }
            K3W_CLK(5);
        };

        int it = 0, s = 0, prev_w = -1, prev_b = 0;
        uint32_t flag_next = (a.flags && w_first < a.W) ? (a.flags[w_first] & SKIP_BITS) : 0u;
        // one pass of the loop = A0 + A1 of walker `w`, then C of the previous live walker; the last pass (w beyond
        // the batch) only drains the pipeline
        for (int w = w_first;; w += gridDim.x, ++it) {
            const bool live = w < a.W;
            bool did_a = false;
            const int b = s & 1;
            if (live) {
                const int cb = it & 1;
                const double* cf = coef_s + (size_t)cb * d.ncoef;
                {   // prefetch the next walker's coefficients (buffer last read two iterations ago, F barriers since)
                    const int wn = w + gridDim.x;
                    if (tid == 0 && wn < a.W) {
                        mbar_expect_tx(&mbar[cb ^ 1], coef_bytes);
                        tma_bulk_g2s(coef_s + (size_t)(cb ^ 1) * d.ncoef, a.coef + (size_t)wn * d.ncoef, coef_bytes,
                                     &mbar[cb ^ 1]);
                    }
                }
                const bool skip = flag_next != 0u;
                {
                    const int wn = w + gridDim.x;
                    flag_next = (a.flags && wn < a.W) ? (a.flags[wn] & SKIP_BITS) : 0u;
                }
                mbar_wait(&mbar[cb], (uint32_t)((it >> 1) & 1));
                if (skip) {
                    bar_sync(BAR_F, KW_NG);      // nobody still polls this mbarrier when thread 0 re-arms it
                } else {
                    did_a = true;
                    double* xs = maps + (size_t)b * map_stride;
                    K3W_CLK_DECL;
                    // ---------------- A0: synthesise the quarter plane (u <= v evaluated, mirrored)
                    {
                        uint32_t r[64];
                        tmem_ld64(tm_mine, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int k = KW_F_FIRST(0); k < KW_SYN; ++k) {
                            const int ez = (int)r[4 * k + 2], ew = (int)r[4 * k + 3];
                            const int sg = ez & 0xffff, u = (ez >> 16) & 0xffff, v = ew & 0xffff;
                            if (u != 0xffff) {
                                const double z = spline_eval(cf, d.nseg, sg, __hiloint2double((int)r[4 * k + 1], (int)r[4 * k]));
                                xs[u * KW_XS + v] = z;
                                xs[v * KW_XS + u] = z;
                            }
                        }
                    }
                    bar_sync(BAR_F, KW_NG);
                    K3W_CLK(0);
                    // ---------------- A1: rows along x, in place, exchanging inside the row pair
                    {
                        double re[16], im[16];
                        for (int base = KW_F_FIRST(gw * 3); base < npair; base += 24) {
                            tmem_ld32(tm_mine + 64, twr);
                            const bool on = lane_on && base + fg < npair;
                            const int u0 = 2 * (base + fg < npair ? base + fg : npair - 1), u1 = u0 + 1;
                            const bool has1 = u1 < H;
                            double2* xbuf = reinterpret_cast<double2*>(xs + u0 * KW_XS);
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int f = fold256(t + 16 * j);
                                double vr = 0.0, vi = 0.0;
                                if (on && f < H) {
                                    vr = xs[u0 * KW_XS + f];
                                    if (has1) vi = xs[u1 * KW_XS + f];
                                }
                                re[j] = vr; im[j] = vi;
                            }
                            __syncwarp();
                            row_pass1(re, im, xbuf, on);
                            __syncwarp();
                            fft256_pass2(t, re, im, xbuf, on);
                            __syncwarp();        // exchange tile read by every lane before the spectra overwrite it
#pragma unroll
                            for (int p = 0; p < 16; ++p) {
                                const int n = t + 16 * rev16(p);
                                if (on && !(n > 128 && t_edge)) {
                                    const int k = fold256(n);
                                    xs[u0 * KW_XS + k] = re[p];
                                    if (has1) xs[u1 * KW_XS + k] = im[p];
                                }
                            }
                        }
                    }
                    __threadfence_block();
                    bar_arrive(BAR_FULL + b, KW_NT);
                    K3W_CLK(1);
                }
            }
            // ---------------- C of the previous live walker (its y convolution ran while this one was transformed)
            if ((did_a || !live) && prev_w >= 0) rows_back(prev_w, prev_b);
            if (!live) break;
            if (did_a) { prev_w = w; prev_b = b; ++s; }
        }
    } else {
        // =====================================================================================  M group
        const int kx = 32 * (gw & 3) + lane, rsel = gw >> 2;
        const int nyq_c = lane >> 3;
        static_assert(KW_NB == 28, "the Nyquist column splits 28 taps into 4 chunks of 7");
        int it = 0, s = 0;
        uint32_t flag_next = (a.flags && w_first < a.W) ? (a.flags[w_first] & SKIP_BITS) : 0u;
        for (int w = w_first; w < a.W; w += gridDim.x, ++it) {
            const bool skip = flag_next != 0u;
            {
                const int wn = w + gridDim.x;
                flag_next = (a.flags && wn < a.W) ? (a.flags[wn] & SKIP_BITS) : 0u;
            }
            if (skip) continue;
            const int b = s & 1;
            double* xs = maps + (size_t)b * map_stride;
            double tap[KW_NB];
            {
                uint32_t r[64];
                tmem_ld64(tm_mine, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < KW_NB; ++j) tap[j] = __hiloint2double((int)r[2 * j + 1], (int)r[2 * j]);
            }
            K3W_CLK_DECL;
            bar_sync(BAR_FULL + b, KW_NT);
            K3W_CLK(6);
            // Nyquist column kx = 128: two rounds of 6 rows per warp, lane = (row slot, chunk of 7 taps)
            double nyq[2];
#pragma unroll
            for (int rd = 0; rd < 2; ++rd) {
                const int nu = (rd * 8 + gw) * 6 + (lane & 7);
                const bool non = KW_M_ON && (lane & 7) < 6 && nu < H;
                double acc = 0.0;
                if (non) {
                    const double* col = xs + 128;
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        const int j = 7 * nyq_c + i, ua = nu - j < 0 ? j - nu : nu - j, ub = nu + j;
                        const double xa = col[ua * KW_XS], xb = j > 0 && ub < H ? col[ub * KW_XS] : 0.0;
                        acc = fma(nyqt[j], xa + xb, acc);
                    }
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                nyq[rd] = acc;
            }
            double acc[KW_UB];
#pragma unroll 1
            for (int pass = KW_M_ON ? 0 : 2; pass < 2; ++pass) {
                const int u0 = (2 * pass + rsel) * KW_UB;
#pragma unroll
                for (int k = 0; k < KW_UB; ++k) acc[k] = 0.0;
                auto taps_of_row = [&](int ii) {
                    const int up = u0 - (KW_NB - 1) + ii, ua = up < 0 ? -up : up;
                    const double x = ua < H ? xs[ua * KW_XS + kx] : 0.0;
#pragma unroll
                    for (int k = 0; k < KW_UB; ++k) {
                        const int j = ii - (KW_NB - 1) - k < 0 ? k + (KW_NB - 1) - ii : ii - (KW_NB - 1) - k;
                        if (j < KW_NB) acc[k] = fma(tap[j], x, acc[k]);
                    }
                };
                // Two straight-line segments with ONE warp-uniform branch between them: the input rows of the second
                // segment lie beyond the map (all zero) for the top row group, a third of its FMAs.  (A test per input
                // row instead keeps the loads from being scheduled ahead of the FMAs: measured 3.04 -> 3.25 ms.)
                constexpr int KW_SP = KW_UB + KW_NB - 3;
#pragma unroll
                for (int ii = 0; ii < KW_SP; ++ii) taps_of_row(ii);
                if (u0 - (KW_NB - 1) + KW_SP < H) {
#pragma unroll
                    for (int ii = KW_SP; ii < KW_UB + 2 * (KW_NB - 1); ++ii) taps_of_row(ii);
                }
                if (pass == 0) {                 // park the first pass in tensor memory (columns 64..111)
                    uint32_t ra[32], rb[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        ra[2 * k] = (uint32_t)__double2loint(acc[k]); ra[2 * k + 1] = (uint32_t)__double2hiint(acc[k]);
                    }
#pragma unroll
                    for (int k = 16; k < KW_UB; ++k) {
                        rb[2 * (k - 16)] = (uint32_t)__double2loint(acc[k]); rb[2 * (k - 16) + 1] = (uint32_t)__double2hiint(acc[k]);
                    }
                    rb[12] = rb[13] = rb[14] = rb[15] = 0u;
                    tmem_st32(tm_mine + 64, ra);
                    tmem_st16(tm_mine + 96, rb);
                }
            }
            K3W_CLK(2);
            tmem_wait_st();
            bar_sync(BAR_M, KW_NG);              // every input has been read: the columns may be overwritten
            {
                const int u1 = (2 + rsel) * KW_UB;
#pragma unroll
                for (int k = 0; k < KW_UB; ++k)
                    if (u1 + k < H) xs[(u1 + k) * KW_XS + kx] = acc[k];
                uint32_t ra[32], rb[16];
                tmem_ld32(tm_mine + 64, ra);
                tmem_ld16(tm_mine + 96, rb);
                tmem_wait_ld();
                const int u0 = rsel * KW_UB;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (u0 + k < H) xs[(u0 + k) * KW_XS + kx] = __hiloint2double((int)ra[2 * k + 1], (int)ra[2 * k]);
#pragma unroll
                for (int k = 16; k < KW_UB; ++k)
                    if (u0 + k < H) xs[(u0 + k) * KW_XS + kx] = __hiloint2double((int)rb[2 * (k - 16) + 1], (int)rb[2 * (k - 16)]);
#pragma unroll
                for (int rd = 0; rd < 2; ++rd) {
                    const int nu = (rd * 8 + gw) * 6 + (lane & 7);
                    if ((lane & 7) < 6 && nu < H && nyq_c == 0) xs[nu * KW_XS + 128] = nyq[rd];
                }
            }
            __threadfence_block();
            bar_arrive(BAR_BDONE + b, KW_NT);
            K3W_CLK(3);
            ++s;
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm_base, 512);
}

}  // namespace

// the warp-specialised kernel serves the geometries the direct form of k3_szmap.cu serves, when two maps fit
bool jx_szmap_ws_ok(const jx_dev& d) {
    if (const char* e = getenv("JX_K3_WS")) if (!atoi(e)) return false;
    if (const char* e = getenv("JX_K3_BFFT")) if (atoi(e)) return false;
    return d.bmix && d.npad == KW_P && d.nbeam <= KW_NB && d.nh <= 4 * KW_UB && d.nh <= 16 * 6 &&
           d.nsynth <= KW_SYN * KW_NG && kw_make_layout(d).total <= 232448;
}

size_t jx_szmap_ws_smem_bytes(const jx_dev& d) { return kw_make_layout(d).total; }

cudaError_t jx_szmap_ws_configure(const jx_dev& d) {
    return cudaFuncSetAttribute(k3w_szmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw_make_layout(d).total);
}

cudaError_t jx_launch_szmap_ws(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                               double* convq, double* tri, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k3_args a;
    a.d = d; a.coef = coef; a.flags = flags; a.W = W; a.convq = convq; a.tri = tri; a.scratch = nullptr; a.scratch2 = nullptr;
    const int grid = W < sm_count ? W : sm_count;
    k3w_szmap_kernel<<<grid, KW_NT, kw_make_layout(d).total, st>>>(a);
    return cudaGetLastError();
}
