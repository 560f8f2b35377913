// K3L -- the map stage for maps whose cyclic convolution length is 512 or 1024 (map sides up to ~1000 pixels):
// same mathematics and phases as k3_szmap.cu (synthesis, row transforms, beam convolution along y, rows back, packed
// triangle out for the filter GEMM; reference joxsz_funcs.py:462-464), but the per-walker quarter-plane
// working set (H x (P/2+1) doubles: 264 KB at N = 255, 1 MB at N = 511) no longer fits the shared memory of an
// SM, so it lives in a per-CTA global scratch that stays resident in the 126 MB L2, and each 16-thread group
// stages one line (row pair or column pair) at a time through shared memory.
//
// Length-P transforms (P = 256 R, R = 2 or 4) are built from the register FFT-256 of jx_fft.cuh by one
// radix-R decimation-in-frequency step:
//
//   y_s[m] = w_P^(s m) sum_{j<R} x[m + 256 j] w_R^(s j),  m = 0..255        X[R k + s] = FFT256(y_s)[k]
//
// Every sequence of the stage is even (x[n] = x[P - n]), so a line is stored as its first P/2+1 samples and
// read through the fold; two real lines ride in the real / imaginary parts of one complex transform.
#include "k3_common.cuh"

namespace {

struct k3l_smem_layout {
    size_t tw, twp, lines, xbuf, coef, mbar, total;
    int lq;     // padded line length (complex elements)
};

__host__ __device__ inline k3l_smem_layout k3l_layout(const jx_dev& d, int nthreads) {
    k3l_smem_layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
    const int groups = nthreads / 16;
    L.lq = d.nq + 1;
    L.tw = take(256 * sizeof(double2));
    L.twp = take((size_t)d.npad * sizeof(double2));
    L.lines = take((size_t)groups * 2 * L.lq * sizeof(double2));       // input + output line per group
    L.xbuf = take((size_t)groups * JX_XB_ELEMS * sizeof(double2));
    L.coef = take((size_t)2 * d.ncoef * sizeof(double));
    L.mbar = take(2 * sizeof(uint64_t));
    L.total = o;
    return L;
}

// (r + i im) *= w_R^(e): e in 0..3 for R = 4 (w_4 = -i), e in 0..1 for R = 2 (w_2 = -1)
template <int R>
JX_D void mul_wr(double& r, double& i, int e) {
    if constexpr (R == 2) {
        if (e & 1) { r = -r; i = -i; }
    } else {
        e &= 3;
        if (e == 1) { double t = r; r = i; i = -t; }          // * (-i)
        else if (e == 2) { r = -r; i = -i; }
        else if (e == 3) { double t = r; r = -i; i = t; }      // * (+i)
    }
}

// Forward DFT of the even sequence x[n] = in[fold(n)], n < P = 256 R: out[K] = X[K] for K <= P/2.
// One 16-thread group; `in` and `out` are distinct shared-memory lines of P/2+1 complex samples.
template <int R>
JX_D void group_fft_even(int t, unsigned gmask, const double2* __restrict__ in, double2* __restrict__ out,
                         const double2* __restrict__ tw256, const double2* __restrict__ twp,
                         double2* __restrict__ xbuf) {
    constexpr int P = 256 * R;
    double re[16], im[16];
#pragma unroll 1
    for (int s = 0; s < R; ++s) {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int m = t + 16 * jj;
            double ar = 0.0, ai = 0.0;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int n = m + 256 * j;
                const double2 v = in[n <= P / 2 ? n : P - n];
                double vr = v.x, vi = v.y;
                mul_wr<R>(vr, vi, s * j);
                ar += vr; ai += vi;
            }
            if (s) {
                const double2 w = twp[(s * m) & (P - 1)];
                const double tr = ar * w.x - ai * w.y;
                ai = ar * w.y + ai * w.x;
                ar = tr;
            }
            re[jj] = ar; im[jj] = ai;
        }
        fft256_pass1(t, re, im, tw256, xbuf);
        __syncwarp(gmask);
        fft256_pass2(t, re, im, xbuf);
        __syncwarp(gmask);
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int K = R * (t + 16 * rev16(p)) + s;
            if (K <= P / 2) out[K] = make_double2(re[p], im[p]);
        }
    }
    __syncwarp(gmask);
}

// Direct y convolution of UB consecutive rows u0 .. u0 + UB - 1 of column kx of the map `in` (row pitch `pitch`, in
// the L2-resident scratch): acc[k] = sum_j tap[|j|] ext(in)[u0 + k - j, kx]; same scheme as k3_szmap.cu's phase B.
constexpr int K3L_NB = JX_BMIX_ROWS, K3L_UB = 32;
JX_D void k3l_yconv(const double* __restrict__ in, int pitch, int kx, int u0, int H, const double (&tap)[K3L_NB],
                    double (&acc)[K3L_UB]) {
    constexpr int NIN = K3L_UB + 2 * (K3L_NB - 1), PF = 8;     // input rows; rows in flight ahead of the arithmetic
    auto fetch = [&](int ii) {
        const int up = u0 - (K3L_NB - 1) + ii, ua = up < 0 ? -up : up;
        return ua < H ? in[(size_t)ua * pitch + kx] : 0.0;
    };
#pragma unroll
    for (int k = 0; k < K3L_UB; ++k) acc[k] = 0.0;
    double xq[PF];
#pragma unroll
    for (int q = 0; q < PF; ++q) xq[q] = fetch(q);
#pragma unroll
    for (int ii = 0; ii < NIN; ++ii) {
        const double x = xq[ii % PF];
        if (ii + PF < NIN) xq[ii % PF] = fetch(ii + PF);
#pragma unroll
        for (int k = 0; k < K3L_UB; ++k) {
            const int j = ii - (K3L_NB - 1) - k < 0 ? k + (K3L_NB - 1) - ii : ii - (K3L_NB - 1) - k;
            if (j < K3L_NB) acc[k] = fma(tap[j], x, acc[k]);
        }
    }
}

template <int R>
__global__ void __launch_bounds__(256, 1) k3l_szmap_kernel(const __grid_constant__ k3_args a) {
    extern __shared__ __align__(128) unsigned char k3l_raw[];
    constexpr int P = 256 * R, Q = P / 2 + 1;
    const jx_dev& d = a.d;
    const int NT = blockDim.x;
    const int H = d.nh, hp8 = d.hp8;
    const k3l_smem_layout L = k3l_layout(d, NT);
    double2* tw_s = reinterpret_cast<double2*>(k3l_raw + L.tw);
    double2* twp_s = reinterpret_cast<double2*>(k3l_raw + L.twp);
    double2* lines = reinterpret_cast<double2*>(k3l_raw + L.lines);
    double2* xbuf_all = reinterpret_cast<double2*>(k3l_raw + L.xbuf);
    double* coef_s = reinterpret_cast<double*>(k3l_raw + L.coef);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(k3l_raw + L.mbar);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = tid >> 4, t = tid & 15, ngroups = NT >> 4;
    const unsigned gmask = 0xffffu << (lane & 16);
    double2* xbuf = xbuf_all + (size_t)grp * JX_XB_ELEMS;
    double2* lin = lines + (size_t)grp * 2 * L.lq;
    double2* lout = lin + L.lq;
    const uint32_t coef_bytes = (uint32_t)(d.ncoef * sizeof(double));
    const int pitch = d.xs_pitch;                               // doubles per row of the scratch map
    double* xs = a.scratch + (size_t)blockIdx.x * hp8 * pitch;
    // direct y convolution (beam of at most 28 samples per side): out of place into a second map, which the later
    // phases then use
    const bool bdirect = a.scratch2 != nullptr;
    double* xc = bdirect ? a.scratch2 + (size_t)blockIdx.x * hp8 * pitch : xs;

    for (int i = tid; i < 256; i += NT) fft256_make_twiddle(i, tw_s[i]);
    for (int i = tid; i < P; i += NT) {
        // exp(-2 pi i m / P), the eighth-turn symmetry is not needed at this size: plain evaluation in double
        const double ang = -2.0 * 3.14159265358979323846 * (double)i / (double)P;
        twp_s[i] = make_double2(cos(ang), sin(ang));
    }
    for (int i = tid; i < hp8 * pitch; i += NT) xs[i] = 0.0;
    if (bdirect)
        for (int i = tid; i < hp8 * pitch; i += NT) xc[i] = 0.0;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int w_first = blockIdx.x;
    if (tid == 0 && w_first < a.W) {
        mbar_expect_tx(&mbar[0], coef_bytes);
        tma_bulk_g2s(coef_s, a.coef + (size_t)w_first * d.ncoef, coef_bytes, &mbar[0]);
    }

    int it = 0;
    for (int w = w_first; w < a.W; w += gridDim.x, ++it) {
        const int buf = it & 1;
        const double* cf = coef_s + (size_t)buf * d.ncoef;
        {
            const int wn = w + gridDim.x;
            if (tid == 0 && wn < a.W) {
                mbar_expect_tx(&mbar[buf ^ 1], coef_bytes);
                tma_bulk_g2s(coef_s + (size_t)(buf ^ 1) * d.ncoef, a.coef + (size_t)wn * d.ncoef, coef_bytes,
                             &mbar[buf ^ 1]);
            }
        }
        // only the bits the profile kernel wrote decide the skip: JX_FLAG_XNONPOS may still be arriving from the
        // X-ray kernel on the side stream and must not split the CTA's control flow (see k3_szmap.cu)
        const bool skip = a.flags && (a.flags[w] & ~(uint32_t)JX_FLAG_XNONPOS) != 0u;
        mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));
        if (skip) {
            __syncthreads();
            continue;
        }

        // ---- A0: synthesise the quarter-plane map into the scratch (u <= v listed, mirrored on store)
        {
            // eight table entries per thread in flight at a time (the table comes from L2)
            const int4* tab = reinterpret_cast<const int4*>(d.synth);
            for (int base = 0; base < d.nsynth; base += 8 * NT) {
                int4 e[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = base + k * NT + tid;
                    e[k] = i < d.nsynth ? __ldg(tab + i) : make_int4(0, 0, 0xffff0000, 0);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int sg = e[k].z & 0xffff, u = (e[k].z >> 16) & 0xffff, v = e[k].w & 0xffff;
                    if (u != 0xffff) {
                        const double z = spline_eval(cf, d.nseg, sg, __hiloint2double(e[k].y, e[k].x));
                        xs[(size_t)u * pitch + v] = z;
                        xs[(size_t)v * pitch + u] = z;
                    }
                }
            }
        }
        __syncthreads();

        // ---- A1: rows along x, in place: xs[u, kx]
        const int npair = (H + 1) >> 1;
        for (int rp = grp; rp < npair; rp += ngroups) {
            const int u0 = 2 * rp, u1 = u0 + 1;
            const bool has1 = u1 < H;
            {
                constexpr int NL = (Q + 15) / 16;            // loads of the whole line in flight together (L2 latency)
                double2 v[NL];
#pragma unroll
                for (int q = 0; q < NL; ++q) {
                    const int f = t + 16 * q;
                    v[q] = make_double2(0.0, 0.0);
                    if (f < H) {
                        v[q].x = xs[(size_t)u0 * pitch + f];
                        if (has1) v[q].y = xs[(size_t)u1 * pitch + f];
                    }
                }
#pragma unroll
                for (int q = 0; q < NL; ++q)
                    if (t + 16 * q < Q) lin[t + 16 * q] = v[q];
            }
            __syncwarp(gmask);
            group_fft_even<R>(t, gmask, lin, lout, tw_s, twp_s, xbuf);
            for (int k = t; k < Q; k += 16) {
                const double2 v = lout[k];
                xs[(size_t)u0 * pitch + k] = v.x;
                if (has1) xs[(size_t)u1 * pitch + k] = v.y;
            }
            __syncwarp(gmask);
        }
        __syncthreads();

        // ---- B: beam convolution along y.  Direct form: 55-tap FMA streams, lane = column, tasks of 32 rows x 32
        // columns dealt to the warps (about a quarter of the FP64 work of the two full-complex column FFTs, and no
        // staging through shared memory)
        if (bdirect) {
            // tasks in column-group-major order, a contiguous share per warp: the taps of a column group are loaded
            // once for the row blocks that follow each other
            const int ncw = (Q + 31) >> 5, nrb = (H + K3L_UB - 1) / K3L_UB, ntask = ncw * nrb, nw = NT >> 5;
            const int t_lo = (int)(((long)warp * ntask) / nw), t_hi = (int)(((long)(warp + 1) * ntask) / nw);
            double tap[K3L_NB];
            int cw_have = -1;
            for (int task = t_lo; task < t_hi; ++task) {
                const int cw = task / nrb, u0 = (task % nrb) * K3L_UB;
                const int kx = 32 * cw + lane;
                const bool on = kx < Q;
                const int kxc = on ? kx : Q - 1;
                if (cw != cw_have) {
#pragma unroll
                    for (int j = 0; j < K3L_NB; ++j) tap[j] = __ldg(d.bmix + (size_t)j * d.bmix_pitch + kxc);
                    cw_have = cw;
                }
                double acc[K3L_UB];
                k3l_yconv(xs, pitch, kxc, u0, H, tap, acc);
#pragma unroll
                for (int k = 0; k < K3L_UB; ++k)
                    if (on && u0 + k < H) xc[(size_t)(u0 + k) * pitch + kx] = acc[k];
            }
        }
        // FFT form: cyclic convolution with the beam spectrum (symmetric: row kx is read), any beam size
        const int ncpair = bdirect ? 0 : (Q + 1) >> 1;
        for (int cp = grp; cp < ncpair; cp += ngroups) {
            const int kx = 2 * cp;
            const bool has1 = kx + 1 < Q;
            for (int f = t; f < Q; f += 16) {
                double2 v = make_double2(0.0, 0.0);
                if (f < H) v = *reinterpret_cast<const double2*>(xs + (size_t)f * pitch + kx);
                lin[f] = v;
            }
            __syncwarp(gmask);
            group_fft_even<R>(t, gmask, lin, lout, tw_s, twp_s, xbuf);
            const double* b0 = d.bhat + (size_t)kx * Q;
            const double* b1 = d.bhat + (size_t)(has1 ? kx + 1 : kx) * Q;
            for (int K = t; K < Q; K += 16) {
                const double2 v = lout[K];
                lin[K] = make_double2(v.x * __ldg(b0 + K), has1 ? v.y * __ldg(b1 + K) : 0.0);
            }
            __syncwarp(gmask);
            group_fft_even<R>(t, gmask, lin, lout, tw_s, twp_s, xbuf);
            for (int u = t; u < H; u += 16)
                *reinterpret_cast<double2*>(xs + (size_t)u * pitch + kx) = lout[u];
            __syncwarp(gmask);
        }
        __syncthreads();

        // ---- C: rows back to pixel space
        for (int rp = grp; rp < npair; rp += ngroups) {
            const int u0 = 2 * rp, u1 = u0 + 1;
            const bool has1 = u1 < H;
            {
                constexpr int NL = (Q + 15) / 16;
                double2 v[NL];
#pragma unroll
                for (int q = 0; q < NL; ++q) {
                    const int f = t + 16 * q;
                    v[q] = make_double2(0.0, 0.0);
                    if (f < Q) {
                        v[q].x = xc[(size_t)u0 * pitch + f];
                        if (has1) v[q].y = xc[(size_t)u1 * pitch + f];
                    }
                }
#pragma unroll
                for (int q = 0; q < NL; ++q)
                    if (t + 16 * q < Q) lin[t + 16 * q] = v[q];
            }
            __syncwarp(gmask);
            group_fft_even<R>(t, gmask, lin, lout, tw_s, twp_s, xbuf);
            // conv_c[u, v] for v >= u goes to the packed triangle of this walker (row u starts at u H - u (u - 1) / 2);
            // the in-place copy only feeds the parity tap
            double* tri0 = a.tri + (size_t)w * d.ktri + (u0 * H - ((u0 * (u0 - 1)) >> 1) - u0);   // + v
            double* tri1 = tri0 + (H - u0 - 1);
            for (int v = t; v < H; v += 16) {
                const double2 o = lout[v];
                if (v >= u0) tri0[v] = o.x;
                if (has1 && v >= u1) tri1[v] = o.y;
                if (a.convq) {
                    xc[(size_t)u0 * pitch + v] = o.x;
                    if (has1) xc[(size_t)u1 * pitch + v] = o.y;
                }
            }
            __syncwarp(gmask);
        }
        __syncthreads();

        if (a.convq) {
            double* cq = a.convq + (size_t)w * H * H;
            for (int i = tid; i < H * H; i += NT) cq[i] = xc[(size_t)(i / H) * pitch + (i % H)];
        }

        __syncthreads();       // the tap's reads of the scratch end before the next walker's phases overwrite it
    }
}

}  // namespace

static int k3l_pick_threads(const jx_dev& d) {
    for (int nt = 256; nt >= 64; nt -= 32)
        if (k3l_layout(d, nt).total <= 232448) return nt;
    return 0;
}

bool jx_szmap_large_supported(const jx_dev& d) { return (d.npad == 512 || d.npad == 1024) && k3l_pick_threads(d) > 0; }

size_t jx_szmap_large_smem_bytes(const jx_dev& d) { return k3l_layout(d, k3l_pick_threads(d)).total; }

cudaError_t jx_szmap_large_configure(const jx_dev& d) {
    const int smem = (int)k3l_layout(d, k3l_pick_threads(d)).total;
    if (d.npad == 512)
        return cudaFuncSetAttribute(k3l_szmap_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return cudaFuncSetAttribute(k3l_szmap_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// scratch: [min(W, sm_count)][hp8][xs_pitch] doubles
cudaError_t jx_launch_szmap_large(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                                  double* convq, double* tri, double* scratch, double* scratch2, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k3_args a;
    a.d = d; a.coef = coef; a.flags = flags; a.W = W; a.convq = convq; a.tri = tri; a.scratch = scratch; a.scratch2 = scratch2;
    const int nt = k3l_pick_threads(d);
    const size_t smem = k3l_layout(d, nt).total;
    const int grid = W < sm_count ? W : sm_count;
    if (d.npad == 512)
        k3l_szmap_kernel<2><<<grid, nt, smem, st>>>(a);
    else
        k3l_szmap_kernel<4><<<grid, nt, smem, st>>>(a);
    return cudaGetLastError();
}
