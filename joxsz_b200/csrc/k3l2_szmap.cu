// K3L2 -- the large-map stage (cyclic length 512 / 1024, direct y convolution).
//
// Same mathematics and phases as k3l_szmap.cu (reference joxsz_funcs.py:462-464) -- results identical at P = 512 and
// equal within rounding at P = 1024 (a quarter of the spectrum samples is taken from its mirror image).  The map of a
// walker (128 x 257 or 256 x 513 doubles) does not fit the shared memory of an SM, so the phases go out of place
// between two per-CTA scratch maps that are reused walker after walker and stay in the 126 MB L2 at 255 pixels
// (A1 xs -> xc, B xc -> xc IN PLACE -- a tile is whole in shared memory before its columns are overwritten, and tiles
// are disjoint column blocks --, C xc -> packed triangle: xs only ever holds the H x H synthesised map, which keeps the
// working set of a CTA at 1.5 instead of 2.1 MB at 511 pixels); one CTA of 256 threads per SM at 255 registers.
//   * A row transform (16-thread group, radix-R decimation in frequency around the register FFT-256) gathers its line
//     ONCE for all branches -- it used to be re-read from L2 per branch, 16 / G dependent chunks of loads each -- and
//     keeps the 32 NS branch inputs in registers, hence the 255-register budget (256 threads: cycles per walker
//     145 k / 830 k at 255 / 511 pixels against 213 k / 1 086 k with 384 threads at 168 registers).
//   * A thread ends a transform with R consecutive samples of the spectrum (branch s of position p holds
//     X[R k + s]; at R = 4 the s = 3 sample comes from lane 15 - t by one shuffle) and stores them as one 16 R-byte run:
//     the scattered 8-byte stores at an 8 R-byte stride cost the 1024-point row transforms a fifth of their time.
//   * The y convolution works on [H rows x 32 columns] tiles of the row spectra that all threads stream into a ring of
//     up to four shared-memory buffers with cp.async (16 bytes per thread and copy; each buffer physically holds the
//     27 mirrored rows, the map rows, 27 rows of zeros and the 28 tap rows of its columns, so the inner loop is loads
//     at immediate offsets and DFMA, no index arithmetic), one __syncthreads per tile; a warp convolves 16-row blocks
//     from shared memory as the 256-point kernels do.  Loads straight from L2 into a register queue were bound by the six scoreboards of a warp
//     (a queue of 16 was no faster than one of 8).  Measured and NOT kept: the same tiles fetched with one
//     cp.async.bulk per 256-byte row (4 096 requests per walker at 511 pixels: request-rate bound, B 318 k -> 438 k
//     cycles), and row pairs of the transforms staged by cp.async.bulk behind the previous FFT (A1 34 k -> 49 k).
//   * The Nyquist column (the 257th / 513th) is convolved apart from shared memory, one output row per lane: as a 33rd
//     tile it cost a tile's work for one active lane.
//   * at P = 1024 the s = 3 branch of the radix-4 decimation is never computed: its spectrum samples are the mirror
//     images X[P - K] of the s = 1 branch (every sequence of the stage is even) (-25 % of the transform work);
//   * The synthesis walks the quarter plane in 32 x 32 tiles (u block <= v block): a warp evaluates rows of a tile and
//     stores them as 256-byte runs, the mirror image goes through a shared-memory transpose and leaves as 256-byte
//     runs too (storing z to [v][u] directly put every lane of a store on its own sector); the next tile's table
//     entries are in flight meanwhile.
//   * the spline coefficients of the next walker arrive by TMA bulk copy while the current walker's transforms and
//     convolution run (they are only read by the synthesis).
// Phase clocks per walker on B200 (scripts/k3_phase_clocks.py, JX_CLK_WORKLOAD=synth255 / synth511), first version ->
// now: 255 pixels A0 13.8 k, A1 48.7 k, B 73.9 k, C 46.4 k = 183 k -> 11.7 / 28.9 / 66.4 / 35.8 = 143 k cycles;
// 511 pixels 66 / 391 / 315 / 311 = 1 083 k -> 43 / 182 / 243 / 190 = 657 k (a few per cent from box to box).
// What bounds B now (0.22 DFMA per clock per scheduler) is documented in profiles/r02_results.md and
// profiles/r02_dfma_latency_yconv_microbench.txt.
#include "k3_common.cuh"

#ifdef JX_K3_CLOCKS
__device__ unsigned long long jx_k3l2_clk[8];
extern "C" int jx_debug_k3l2_clocks(unsigned long long* out8) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, jx_k3l2_clk, sizeof(jx_k3l2_clk));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(jx_k3l2_clk, z, sizeof(z));
    return e == cudaSuccess ? 0 : -1;
}
#define K3M_CLK_DECL long long k3_t0 = clock64()
#define K3M_CLK(i) do { if (threadIdx.x == 0) { long long k3_t1 = clock64(); atomicAdd(&jx_k3l2_clk[i], (unsigned long long)(k3_t1 - k3_t0)); k3_t0 = k3_t1; } } while (0)
#else
#define K3M_CLK_DECL
#define K3M_CLK(i)
#endif

// Timing experiments of phase B (scripts/k3_phase_clocks.py with JX_CLK_DEFS; results of such builds are wrong):
// -DK3M_NO_STG drops the result stores, -DK3M_NO_FETCH the tile fetches after the first ring fill.
#ifdef K3M_NO_STG
#define K3M_EXP_STORE(c) ((c) && a.W < 0)
#else
#define K3M_EXP_STORE(c) (c)
#endif
#ifdef K3M_NO_FETCH
#define K3M_EXP_FETCH(c) ((c) && a.W < 0)
#else
#define K3M_EXP_FETCH(c) (c)
#endif

namespace {

constexpr int K3M_NT_A = 512, K3M_NT_B = 384, K3M_NT_C = 256;       // threads per CTA (one CTA per SM): 16 warps at 128 registers, or 12 at 168
#ifndef K3M_UB_N
#define K3M_UB_N 16
#endif
constexpr int K3M_NB = JX_BMIX_ROWS, K3M_UB = K3M_UB_N;      // taps, rows per convolution task
#ifndef K3M_GATHER_N
#define K3M_GATHER_N 32
#endif
// samples per gather chunk of a row transform (x 2 rows = loads in flight per thread; all of a chunk are issued, then
// all are consumed, so they share scoreboards without harm).  Cycles per walker in A1 at 255 / 511 pixels: 8 samples
// 32.8 k / 206 k, 16 samples 30.4 k / 192 k, 32 samples 28.9 k / 182 k
constexpr int K3M_GATHER = K3M_GATHER_N;
#ifndef K3M_TS_MIN_N
#define K3M_TS_MIN_N 2
#endif
// tiles per step (= per block barrier) of the y convolution, at least, when the ring holds two such steps: at 255 pixels
// two tiles per step, i.e. two blocks per warp between barriers: B 64.8 k -> 63.1 k cycles per walker
constexpr int K3M_TS_MIN = K3M_TS_MIN_N;
constexpr int K3M_TILES = 4;                                 // tile buffers of the y convolution, at most

// a convolution tile of 32 columns: the 27 mirrored rows -27 .. -1 (copies of rows 27 .. 1), the H rows of the map, 27
// rows of zeros (the rows beyond the map), then the 28 tap rows of its columns.  Every input of every output row is
// then physically present at tile row (27 + its row): the inner loop loads with immediate offsets and no index
// arithmetic (|row|, row < H ? ... cost six integer instructions per load, 45 % of the loop)
constexpr int K3M_PAD = K3M_NB - 1;
__host__ __device__ constexpr int k3m_tile_rows(int H) { return H + 2 * K3M_PAD + K3M_NB; }

struct k3m_layout {
    size_t tw, twp, coef, nyq, mbar, xbuf, total;
    int ntile;               // [H][32] tiles of the y convolution that fit the arena (exchange tiles and what is left)
};

__host__ __device__ inline k3m_layout k3m_make_layout(const jx_dev& d, int nthreads) {
    k3m_layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
    L.tw = take(256 * sizeof(double2));
    L.twp = take((size_t)(d.npad / 2) * sizeof(double2));
    L.coef = take((size_t)d.ncoef * sizeof(double));
    L.nyq = take((size_t)(d.nh + 1 + K3M_NB) * sizeof(double));      // Nyquist column, a zero, its taps
    L.mbar = take(sizeof(uint64_t));
    o = (o + 127) & ~size_t(127);
    // the arena: exchange tiles of the transform groups in A1 / C, ring of convolution tiles in B
    L.xbuf = take((size_t)(nthreads / 16) * JX_XB_ELEMS * sizeof(double2));
    const size_t tile = (size_t)k3m_tile_rows(d.nh) * 32 * sizeof(double), cap = 232448;
    size_t n = cap > L.xbuf ? (cap - L.xbuf) / tile : 0;
    if (n > K3M_TILES) n = K3M_TILES;
    L.ntile = (int)n;
    if (L.xbuf + n * tile > o) o = L.xbuf + n * tile;
    L.total = o;
    return L;
}

template <int R>
JX_D void mul_wr2(double& r, double& i, int e) {
    if constexpr (R == 2) {
        if (e & 1) { r = -r; i = -i; }
    } else {
        e &= 3;
        if (e == 1) { double t = r; r = i; i = -t; }          // * (-i)
        else if (e == 2) { r = -r; i = -i; }
        else if (e == 3) { double t = r; r = -i; i = t; }      // * (+i)
    }
}

// Forward DFT of the even sequence x[n] = ld(fold(n)), n < P = 256 R, by one 16-thread group: st(K, re, im) receives
// X[K] for every K <= P/2 exactly once.  Radix-R decimation in frequency around the register FFT-256:
//   y_s[m] = w_P^(s m) sum_j x[m + 256 j] w_R^(s j),   X[R k + s] = FFT256(y_s)[k];
// for R = 4 the branch s = 3 is skipped: X[4 k + 3] = X[P - 4 k - 3] = X[4 (255 - k) + 1] comes out of branch s = 1.
template <int R, class LD, class STV, class ST1>
JX_D void k3m_group_fft_even(int t, unsigned gmask, LD&& ld, STV&& stv, ST1&& st1, const double2* __restrict__ tw256,
                             const double2* __restrict__ twp, double2* __restrict__ xbuf) {
    constexpr int P = 256 * R, NS = R == 4 ? 3 : R;
    // The line is gathered ONCE for all branches (it used to be re-read from L2 per branch: the load latency of
    // 16 / G dependent chunks per branch was most of a row transform): 32 NS registers hold the branch inputs, then
    // the register FFT-256 runs branch after branch.  Same operations in the same order per output as before.
    double re[NS][16], im[NS][16];
    // the gather runs in chunks of G samples: all loads of a chunk are in flight together (L2 latency), and the
    // compiler barrier between chunks keeps it from hoisting every load of the line
    constexpr int G = K3M_GATHER / R;
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += G) {
        double2 v[G][R];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int n = t + 16 * (j0 + g) + 256 * j;
                v[g][j] = ld(n <= P / 2 ? n : P - n);
            }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int m = t + 16 * (j0 + g);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                double ar = 0.0, ai = 0.0;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    double vr = v[g][j].x, vi = v[g][j].y;
                    mul_wr2<R>(vr, vi, s * j);
                    ar += vr; ai += vi;
                }
                if (s) {
                    const double2 w = twp[s * m];
                    const double tr = ar * w.x - ai * w.y;
                    ai = ar * w.y + ai * w.x;
                    ar = tr;
                }
                re[s][j0 + g] = ar; im[s][j0 + g] = ai;
            }
        }
        asm volatile("" ::: "memory");
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        fft256_pass1(t, re[s], im[s], tw256, xbuf);
        __syncwarp(gmask);
        fft256_pass2(t, re[s], im[s], xbuf);
        __syncwarp(gmask);
    }
    // Position p of branch s holds X[R k + s], k = t + 16 rev16(p): a thread owns R consecutive samples of the
    // spectrum (for R = 4 the s = 3 sample X[4 k + 3] = X[P - 4 k - 3] = X[4 (255 - k) + 1] sits at position 15 - p of
    // branch 1 in lane 15 - t: one shuffle), and stores them as one 16 R-byte run instead of R scattered 8-byte
    // stores at a 8 R-byte stride (which cost the row transforms of the 1024-point maps a third of their time).
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        double vr[R], vi[R];
#pragma unroll
        for (int s = 0; s < NS; ++s) { vr[s] = re[s][p]; vi[s] = im[s][p]; }
        if constexpr (R == 4) {
            vr[3] = __shfl_xor_sync(gmask, re[1][15 - p], 15);
            vi[3] = __shfl_xor_sync(gmask, im[1][15 - p], 15);
        }
        const int k = t + 16 * rev16(p);
        if (k < 128) stv(R * k, vr, vi);
        else if (k == 128) st1(P / 2, vr[0], vi[0]);
    }
}

JX_D void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
JX_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
JX_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// y convolution of UB consecutive rows (from u0) of one column of a shared-memory tile (lane = column): input row
// u0 - 27 + ii sits at tile row u0 + ii.  Same scheme and summation order as k3l_szmap.cu / k3_szmap.cu
JX_D void k3m_yconv(const double* __restrict__ tile, int lane, int u0, const double (&tap)[K3M_NB], double (&acc)[K3M_UB]) {
    constexpr int NIN = K3M_UB + 2 * K3M_PAD;
    const double* in = tile + u0 * 32 + lane;
#pragma unroll
    for (int k = 0; k < K3M_UB; ++k) acc[k] = 0.0;
#pragma unroll
    for (int ii = 0; ii < NIN; ++ii) {
        const double x = in[ii * 32];
#pragma unroll
        for (int k = 0; k < K3M_UB; ++k) {
            const int j = ii - K3M_PAD - k < 0 ? k + K3M_PAD - ii : ii - K3M_PAD - k;
            if (j < K3M_NB) acc[k] = fma(tap[j], x, acc[k]);
        }
    }
}

template <int R, int K3M_NT>
__global__ void __launch_bounds__(K3M_NT, 1) k3l2_szmap_kernel(const __grid_constant__ k3_args a) {
    extern __shared__ __align__(128) unsigned char k3m_raw[];
    constexpr int P = 256 * R, Q = P / 2 + 1, NT = K3M_NT;
    const jx_dev& d = a.d;
    const int H = d.nh, hp8 = d.hp8;
    const k3m_layout L = k3m_make_layout(d, K3M_NT);
    double2* tw_s = reinterpret_cast<double2*>(k3m_raw + L.tw);
    double2* twp_s = reinterpret_cast<double2*>(k3m_raw + L.twp);
    double2* xbuf_all = reinterpret_cast<double2*>(k3m_raw + L.xbuf);
    double* coef_s = reinterpret_cast<double*>(k3m_raw + L.coef);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(k3m_raw + L.mbar);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 4, t = tid & 15, ngroups = NT >> 4;
    const unsigned gmask = 0xffffu << (lane & 16);
    double2* xbuf = xbuf_all + (size_t)grp * JX_XB_ELEMS;
    const uint32_t coef_bytes = (uint32_t)(d.ncoef * sizeof(double));
    const int pitch = d.xs_pitch;
    double* xs = a.scratch + (size_t)blockIdx.x * hp8 * pitch;      // synthesised map (H x H)
    double* xc = a.scratch2 + (size_t)blockIdx.x * hp8 * pitch;     // row spectra, convolved in place

    for (int i = tid; i < 256; i += NT) fft256_make_twiddle(i, tw_s[i]);
    for (int i = tid; i < P / 2; i += NT) {          // s m < P / 2 for the branches that are computed: half a turn
        const double ang = -2.0 * 3.14159265358979323846 * (double)i / (double)P;
        twp_s[i] = make_double2(cos(ang), sin(ang));
    }
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int w_first = blockIdx.x;
    if (tid == 0 && w_first < a.W) {
        mbar_expect_tx(mbar, coef_bytes);
        tma_bulk_g2s(coef_s, a.coef + (size_t)w_first * d.ncoef, coef_bytes, mbar);
    }

    const int npair = (H + 1) >> 1;
    int it = 0;
    for (int w = w_first; w < a.W; w += gridDim.x, ++it) {
        // only the bits the profile kernel wrote decide the skip (see k3_szmap.cu)
        const bool skip = a.flags && (a.flags[w] & ~(uint32_t)JX_FLAG_XNONPOS) != 0u;
        K3M_CLK_DECL;
        mbar_wait(mbar, (uint32_t)(it & 1));
        if (!skip) {
            // ---- A0: synthesise the quarter-plane map, 32 x 32 tiles with u block <= v block.  A warp evaluates rows of
            // the tile and stores them as 256-byte runs; the mirror image goes through a shared-memory transpose and
            // leaves as 256-byte runs too (storing z to [v][u] straight away put every lane of a store on its own
            // sector: the synthesis of a 511-pixel map took 63-80 k cycles, 9-11 % of the kernel)
            const int4* tab = reinterpret_cast<const int4*>(d.synth_tiles);
            double* tt_all = reinterpret_cast<double*>(k3m_raw + L.xbuf);      // two [32][33] transpose tiles in the arena
            const int nt = (H + 31) >> 5, nwarp = NT >> 5;
            constexpr int RPW = (32 + (NT >> 5) - 1) / (NT >> 5);              // tile rows per warp
            int4 e[RPW], en[RPW];
            auto tile_load = [&](int tile, int4 (&x)[RPW]) {
#pragma unroll
                for (int q = 0; q < RPW; ++q) {
                    const int i = warp + q * nwarp;
                    x[q] = i < 32 && tile < d.nsynth_tiles ? __ldg(tab + (size_t)tile * 1024 + i * 32 + lane)
                                                          : make_int4(0, 0, 0xffff0000, 0);
                }
            };
            tile_load(0, e);
            int tile = 0;
            for (int ub = 0; ub < nt; ++ub)
                for (int vb = ub; vb < nt; ++vb, ++tile) {
                    tile_load(tile + 1, en);                                   // the next tile's entries are in flight
                    double* tt = tt_all + (tile & 1) * (32 * 33);
#pragma unroll
                    for (int q = 0; q < RPW; ++q) {
                        const int i = warp + q * nwarp;
                        const int sg = e[q].z & 0xffff, u = (e[q].z >> 16) & 0xffff, v = e[q].w & 0xffff;
                        double z = 0.0;
                        if (u != 0xffff) {
                            z = spline_eval(coef_s, d.nseg, sg, __hiloint2double(e[q].y, e[q].x));
                            __stcg(xs + (size_t)u * pitch + v, z);
                        }
                        if (i < 32) tt[i * 33 + lane] = z;
                    }
                    __syncthreads();         // also orders this tile's reads of `tt` before the writes two tiles on
                    if (ub != vb) {
#pragma unroll
                        for (int q = 0; q < RPW; ++q) {
                            const int i = warp + q * nwarp;                    // row of the mirrored tile
                            const int u = 32 * vb + i, v = 32 * ub + lane;
                            if (i < 32 && u < H && v < H) __stcg(xs + (size_t)u * pitch + v, tt[lane * 33 + i]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < RPW; ++q) e[q] = en[q];
                }
        }
        __syncthreads();                       // the coefficients have been read (or are not needed): refill them
        {
            const int wn = w + gridDim.x;
            if (tid == 0 && wn < a.W) {
                mbar_expect_tx(mbar, coef_bytes);
                tma_bulk_g2s(coef_s, a.coef + (size_t)wn * d.ncoef, coef_bytes, mbar);
            }
        }
        if (skip) continue;
        K3M_CLK(0);

        // ---- A1: rows along x, xs -> xc
        for (int rp = grp; rp < npair; rp += ngroups) {
            const int u0 = 2 * rp, u1 = u0 + 1;
            const bool has1 = u1 < H;
            const double* r0 = xs + (size_t)u0 * pitch;
            const double* r1 = xs + (size_t)(has1 ? u1 : u0) * pitch;
            double* o0 = xc + (size_t)u0 * pitch;
            double* o1 = xc + (size_t)u1 * pitch;
            k3m_group_fft_even<R>(
                t, gmask,
                [&](int f) { return f < H ? make_double2(__ldcg(r0 + f), has1 ? __ldcg(r1 + f) : 0.0) : make_double2(0.0, 0.0); },
                [&](int K0, const double (&vr)[R], const double (&vi)[R]) {
#pragma unroll
                    for (int i = 0; i < R; i += 2) {
                        __stcg(reinterpret_cast<double2*>(o0 + K0 + i), make_double2(vr[i], vr[i + 1]));
                        if (has1) __stcg(reinterpret_cast<double2*>(o1 + K0 + i), make_double2(vi[i], vi[i + 1]));
                    }
                },
                [&](int K, double vr, double vi) { __stcg(o0 + K, vr); if (has1) __stcg(o1 + K, vi); },
                tw_s, twp_s, xbuf);
        }
        __syncthreads();
        K3M_CLK(1);

        // ---- B: 55-tap convolution along y, xc -> xc in place.  [H rows x 32 columns] tiles of the row spectra stream through a
        // ring of shared-memory buffers (cp.async, 16 bytes per thread and copy: a register prefetch queue is never
        // deeper than the six scoreboards of a warp); lane = column, a warp takes a block of K3M_UB rows of a tile.
        // Rows per block: 16 (K3M_UB).  32-row blocks with two tiles per step (so that every warp still has a block
        // at 255 pixels) were measured equal (B 62.4 k cycles per walker either way at 255 pixels, 228 k -> 252 k at 511):
        // the step logic below stays general, the default stays 16.
        {
            // the Q - 1 = P / 2 columns below the Nyquist frequency make whole 32-column tiles; the Nyquist column
            // is done with lane = row (as a tile it would cost a tile's work for one active lane)
            const int ncw = (Q - 1) >> 5, nrb = (H + K3M_UB - 1) / K3M_UB, nw = NT >> 5;
            const int ntb = L.ntile, tile_elems = k3m_tile_rows(H) * 32;
            int ts = nrb < nw ? nw / nrb : 1;                    // tiles per step
            if (ts < K3M_TS_MIN) ts = K3M_TS_MIN;
            if (ntb / ts < 2) ts = 1;
            const int nslot = ntb / ts, nstep = (ncw + ts - 1) / ts;      // ring slots of ts tiles each; steps
            double* tiles = reinterpret_cast<double*>(k3m_raw + L.xbuf);
            auto tile_fetch = [&](int cw, int buf) {
                double* dst = tiles + (size_t)buf * tile_elems;
                const double* src = xc + 32 * cw;
                for (int i = tid; i < (H + K3M_PAD) * 16; i += NT) {          // tile rows 0 .. 26 + H <- rows 27 .. 1, 0 .. H - 1
                    const int tr = i >> 4, c = (i & 15) * 2;
                    int r = tr - K3M_PAD;
                    r = r < 0 ? -r : r;
                    if (r < H) cp_async16(dst + tr * 32 + c, src + (size_t)r * pitch + c);
                }
                const double* tsrc = d.bmix + 32 * cw;
                for (int i = tid; i < K3M_NB * 16; i += NT) {
                    const int r = i >> 4, c = (i & 15) * 2;
                    cp_async16(dst + (H + 2 * K3M_PAD + r) * 32 + c, tsrc + (size_t)r * d.bmix_pitch + c);
                }
            };
            auto step_fetch = [&](int st) {
                for (int i = 0; i < ts; ++i)
                    if (st * ts + i < ncw) tile_fetch(st * ts + i, (st % nslot) * ts + i);
            };
            for (int i = tid; i < K3M_PAD * 32 * ntb; i += NT)           // the rows beyond the map
                tiles[(size_t)(i / (K3M_PAD * 32)) * tile_elems + (K3M_PAD + H) * 32 + i % (K3M_PAD * 32)] = 0.0;
            for (int i = 0; i < nslot - 1; ++i) {        // a group per ring slot, empty or not: uniform counting
                if (i < nstep) step_fetch(i);
                cp_async_commit();
            }
            // the Nyquist column, a zero behind it and its taps go to shared memory by plain loads
            double* nyq_s = reinterpret_cast<double*>(k3m_raw + L.nyq);
            for (int i = tid; i <= H + K3M_NB; i += NT)
                nyq_s[i] = i < H ? __ldcg(xc + (size_t)i * pitch + (Q - 1))
                         : i == H ? 0.0 : __ldg(d.bmix + (size_t)(i - H - 1) * d.bmix_pitch + (Q - 1));
            for (int st = 0; st < nstep; ++st) {
                // steps st + 1 .. st + nslot - 2 may still be in flight; one barrier per step: past it the tiles of
                // step st have landed for every thread and every warp is done with step st - 1, whose slot the next
                // fetch refills
                if (nslot >= 4) cp_async_wait<2>(); else if (nslot == 3) cp_async_wait<1>(); else cp_async_wait<0>();
                __syncthreads();
                if (K3M_EXP_FETCH(st + nslot - 1 < nstep)) step_fetch(st + nslot - 1);
                cp_async_commit();
                if (st == 0) {
                    // Nyquist column: one output row per lane (rows dealt over the warps), inputs and taps in the
                    // order of k3m_yconv
                    for (int r = warp + nw * lane; r < H; r += nw * 32) {
                        double acc = 0.0;
#pragma unroll
                        for (int dd = -(K3M_NB - 1); dd <= K3M_NB - 1; ++dd) {
                            const int up = r + dd, ua = up < 0 ? -up : up;
                            acc = fma(nyq_s[H + 1 + (dd < 0 ? -dd : dd)], nyq_s[ua < H ? ua : H], acc);
                        }
                        __stcg(xc + (size_t)r * pitch + (Q - 1), acc);
                    }
                }
                for (int task = warp; task < ts * nrb; task += nw) {
                    const int ti = task / nrb, u0 = (task - ti * nrb) * K3M_UB, cw = st * ts + ti;
                    if (cw >= ncw) continue;
                    const int kx = 32 * cw + lane;
                    const double* tile = tiles + (size_t)((st % nslot) * ts + ti) * tile_elems;
                    double tap[K3M_NB];
#pragma unroll
                    for (int j = 0; j < K3M_NB; ++j) tap[j] = tile[(H + 2 * K3M_PAD + j) * 32 + lane];
                    double acc[K3M_UB];
                    k3m_yconv(tile, lane, u0, tap, acc);
#pragma unroll
                    for (int k = 0; k < K3M_UB; ++k)
                        if (K3M_EXP_STORE(u0 + k < H)) __stcg(xc + (size_t)(u0 + k) * pitch + kx, acc[k]);
                }
            }
            cp_async_wait<0>();
        }
        __syncthreads();
        K3M_CLK(2);

        // ---- C: rows back to pixel space, xc -> packed triangle (and the quarter-plane tap)
        double* cq = a.convq ? a.convq + (size_t)w * H * H : nullptr;
        for (int rp = grp; rp < npair; rp += ngroups) {
            const int u0 = 2 * rp, u1 = u0 + 1;
            const bool has1 = u1 < H;
            const double* r0 = xc + (size_t)u0 * pitch;
            const double* r1 = xc + (size_t)(has1 ? u1 : u0) * pitch;
            double* tri0 = a.tri + (size_t)w * d.ktri + (u0 * H - ((u0 * (u0 - 1)) >> 1) - u0);   // + v
            double* tri1 = tri0 + (H - u0 - 1);
            k3m_group_fft_even<R>(
                t, gmask,
                [&](int f) { return make_double2(__ldcg(r0 + f), has1 ? __ldcg(r1 + f) : 0.0); },
                [&](int v0, const double (&vr)[R], const double (&vi)[R]) {
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int v = v0 + i;
                        if (v < H) {
                            if (v >= u0) tri0[v] = vr[i];
                            if (has1 && v >= u1) tri1[v] = vi[i];
                            if (cq) {
                                cq[(size_t)u0 * H + v] = vr[i];
                                if (has1) cq[(size_t)u1 * H + v] = vi[i];
                            }
                        }
                    }
                },
                [&](int v, double vr, double vi) {
                    if (v < H) {                 // never with a real beam (P >= 2 H - 1 + beam - 1), kept for safety
                        if (v >= u0) tri0[v] = vr;
                        if (has1 && v >= u1) tri1[v] = vi;
                        if (cq) { cq[(size_t)u0 * H + v] = vr; if (has1) cq[(size_t)u1 * H + v] = vi; }
                    }
                },
                tw_s, twp_s, xbuf);
        }
        __syncthreads();        // the exchange tiles are free: the next synthesis transposes through the same arena
        K3M_CLK(3);
    }
}

}  // namespace

static int k3m_threads() {
    // measured on B200 (ms per 4 096-walker launch at 255 / 511 pixels) before the line gather was merged over the
    // branches: 512 threads 2.73 / 15.33, 384 threads 2.62 / 15.13, 256 threads 2.69 / 15.62.  With 32 NS live
    // registers of branch inputs the transforms want the 255-register budget: 256 threads are the default now
    // (cycles per walker at 255 / 511 pixels: 145 k / 830 k against 213 k / 1 086 k with 384 threads)
    if (const char* e = getenv("JX_K3L2_NT")) {
        if (atoi(e) == K3M_NT_A) return K3M_NT_A;
        if (atoi(e) == K3M_NT_B) return K3M_NT_B;
    }
    return K3M_NT_C;
}

// Measured on B200 against k3l_szmap_kernel (8 192 walkers, per 4 096-walker launch): 511 pixels 19.6 -> 15.1 ms,
// 255 pixels 2.70 -> 2.62 ms.  JX_K3L2=0 selects the older kernel.
bool jx_szmap_large2_ok(const jx_dev& d) {
    if (const char* e = getenv("JX_K3L2")) if (!atoi(e)) return false;
    return (d.npad == 512 || d.npad == 1024) && d.bmix && d.nbeam <= K3M_NB && k3m_make_layout(d, k3m_threads()).total <= 232448 &&
           k3m_make_layout(d, k3m_threads()).ntile >= 2;
}

size_t jx_szmap_large2_smem_bytes(const jx_dev& d) { return k3m_make_layout(d, k3m_threads()).total; }

template <class F>
static auto k3m_dispatch(const jx_dev& d, F&& f) {
    const int nt = k3m_threads();
    if (d.npad == 512)
        return nt == K3M_NT_A ? f(k3l2_szmap_kernel<2, K3M_NT_A>, nt)
             : nt == K3M_NT_B ? f(k3l2_szmap_kernel<2, K3M_NT_B>, nt) : f(k3l2_szmap_kernel<2, K3M_NT_C>, nt);
    return nt == K3M_NT_A ? f(k3l2_szmap_kernel<4, K3M_NT_A>, nt)
         : nt == K3M_NT_B ? f(k3l2_szmap_kernel<4, K3M_NT_B>, nt) : f(k3l2_szmap_kernel<4, K3M_NT_C>, nt);
}

cudaError_t jx_szmap_large2_configure(const jx_dev& d) {
    return k3m_dispatch(d, [&](auto kern, int nt) {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3m_make_layout(d, nt).total);
    });
}

// scratch / scratch2: [min(W, sm_count)][hp8][xs_pitch] doubles each
cudaError_t jx_launch_szmap_large2(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                                   double* convq, double* tri, double* scratch, double* scratch2, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k3_args a;
    a.d = d; a.coef = coef; a.flags = flags; a.W = W; a.convq = convq; a.tri = tri; a.scratch = scratch; a.scratch2 = scratch2;
    const int grid = W < sm_count ? W : sm_count;
    return k3m_dispatch(d, [&](auto kern, int nt) {
        kern<<<grid, nt, k3m_make_layout(d, nt).total, st>>>(a);
        return cudaGetLastError();
    });
}
