// K4 -- X-ray band emissivity from the cooling-function tables, annulus projection and the Cash
// statistic, one warp per walker.
//
// Replaces Fit.calcProfiles -> Band.calcProjProfile -> CountRate.getCountRate (mbproj2; call site
// joxsz_funcs.py:527), the positivity gate (:529-532) and mylikeFromProfs / cashLogLikelihood (:495-505):
//
//   lnT    = ln(clip(T_X, Tmin, Tmax))
//   rate_s = (exp(interp(lnT, Tlog, t0)) + (exp(interp(lnT, Tlog, t1)) - exp(interp(lnT, Tlog, t0))) * Z) * ne^2
//   pred_j = (sum_s projvols[j, s] rate_s) * areascale_j * exposure_j + bkg_j * backscale
//   like   = sum_bands [ sum_j cts_j ln pred_j - sum_j pred_j ]   over bins with cts_j not NaN
//
// Lane a owns shell a for the emissivity and annulus a for the projection; the per-walker rate vector
// goes through shared memory.  With at most 16 annuli (the shipped layout has 15) a warp carries two walkers,
// one per half-warp: the kernel is bound by the latency of its dependent table look-ups, not by throughput.  Tables (2 x nb x ntab doubles) and projvols are read through the
// read-only path and stay L1/L2 resident (every warp reads the same 16 KB).
#include "jx_common.cuh"

namespace {

constexpr int K4_WARPS = 8;

struct k4_args {
    jx_dev d;
    const double *theta, *ne_ann, *tx_ann;
    int W;
    double *pred, *cash;
    uint32_t* flags;
};

// numpy.interp on an increasing grid, split in two so that the search runs once per shell and is shared by
// the 2 x nb tables: `np_interp_locate` returns the segment index (or -1 / -2 / -3 for "clamp to the first
// value" / "clamp to the last value" / NaN abscissa) and np_interp_eval applies numpy's formula
// slope * (x - x0) + f0 on that segment.
JX_D int np_interp_locate(double x, const double* __restrict__ xp, int n) {
    if (x != x) return -3;
    if (x > __ldg(xp + n - 1)) return -2;
    if (x < __ldg(xp)) return -1;
    if (x == __ldg(xp + n - 1)) return -2;
    int lo = 0, hi = n - 1;          // invariant: xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x >= __ldg(xp + mid)) lo = mid; else hi = mid;
    }
    return lo;
}

JX_D double np_interp_eval(int lo, double x, double x0, const double* __restrict__ xp,
                           const double* __restrict__ fp, int n) {
    if (lo == -3) return x;
    if (lo == -2) return __ldg(fp + n - 1);
    if (lo == -1) return __ldg(fp);
    const double f0 = __ldg(fp + lo);
    const double slope = (__ldg(fp + lo + 1) - f0) / (__ldg(xp + lo + 1) - x0);
    return slope * (x - x0) + f0;
}

// sum over the lanes of a walker: the whole warp (PER_WARP = 1) or its half-warp (PER_WARP = 2); the xor tree over
// 16 lanes gives bit-identical sums to the 32-lane tree with zeros in the upper half
template <int PER_WARP>
JX_D double k4_sum(double v) {
#pragma unroll
    for (int o = 16 / PER_WARP; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int PER_WARP>
__global__ void __launch_bounds__(K4_WARPS * 32) k4_xray_kernel(const __grid_constant__ k4_args a) {
    extern __shared__ double k4_smem[];
    const jx_dev& d = a.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int LANES = 32 / PER_WARP;
    const int sub = lane / LANES;                                   // which walker of the warp
    const int w0 = (blockIdx.x * K4_WARPS + warp) * PER_WARP;
    if (w0 >= a.W) return;                                          // whole warps leave
    const bool live = w0 + sub < a.W;                               // an odd batch leaves the last half-warp idle
    const int w = live ? w0 + sub : w0;
    double* rate_s = k4_smem + (size_t)(warp * PER_WARP + sub) * d.na;

    const int zsrc = d.slot_src[JX_ZMET], bsrc = d.slot_src[JX_BACKSCALE];
    const double Z = zsrc < 0 ? d.slot_val[JX_ZMET] : a.theta[(size_t)w * d.ndim + zsrc];
    const double backscale = bsrc < 0 ? d.slot_val[JX_BACKSCALE] : a.theta[(size_t)w * d.ndim + bsrc];

    double like = 0.0;
    bool nonpos = false;
    // one lane per shell/annulus: na <= 32 is enforced by jx_create
    const int s = lane % LANES;         // shell / annulus index of this lane
    double ne = 0.0, lnT = 0.0;
    if (s < d.na) {
        ne = a.ne_ann[(size_t)w * d.na + s];
        double T = a.tx_ann[(size_t)w * d.na + s];
        // np.clip propagates NaN
        double Tc = (T != T) ? T : fmin(fmax(T, d.tmin), d.tmax);
        lnT = log(Tc);
    }
    int seg_t = -3;
    double x0_t = 0.0;
    if (s < d.na) {
        seg_t = np_interp_locate(lnT, d.tlog, d.ntab);
        if (seg_t >= 0) x0_t = __ldg(d.tlog + seg_t);
    }
    for (int b = 0; b < d.nb; ++b) {
        if (s < d.na) {
            double r0 = exp(np_interp_eval(seg_t, lnT, x0_t, d.tlog, d.lnrate0 + (size_t)b * d.ntab, d.ntab));
            double r1 = exp(np_interp_eval(seg_t, lnT, x0_t, d.tlog, d.lnrate1 + (size_t)b * d.ntab, d.ntab));
            rate_s[s] = (r0 + (r1 - r0) * Z) * (ne * ne);
        }
        __syncwarp();
        double t1 = 0.0, t2 = 0.0;
        if (s < d.na) {
            const double* pv = d.projvols + (size_t)s * d.na;
            double proj = 0.0;
            for (int k = 0; k < d.na; ++k) proj += __ldg(pv + k) * rate_s[k];
            size_t o = (size_t)b * d.na + s;
            double pred = proj * __ldg(d.srcscale + o) + __ldg(d.bkgterm + o) * backscale;
            if (a.pred && live) a.pred[((size_t)w * d.nb + b) * d.na + s] = pred;
            if (!(pred > 0.0)) nonpos = true;
            double c = __ldg(d.cts + o);
            if (c == c) { t1 = c * log(pred); t2 = pred; }
        }
        __syncwarp();
        double lb = k4_sum<PER_WARP>(t1) - k4_sum<PER_WARP>(t2);
        like += isfinite(lb) ? lb : jx_neg_inf();
    }
    const unsigned np_mask = __ballot_sync(0xffffffffu, nonpos);
    nonpos = (np_mask & (PER_WARP == 1 ? 0xffffffffu : (0xffffu << (16 * sub)))) != 0u;
    if (s == 0 && live) {
        if (a.cash) a.cash[w] = nonpos ? jx_neg_inf() : like;
        if (a.flags && nonpos) a.flags[w] |= JX_FLAG_XNONPOS;
    }
}

__global__ void __launch_bounds__(K4_WARPS * 32)
k4_cash_kernel(jx_dev d, const double* __restrict__ pred, int W, double* __restrict__ cash) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * K4_WARPS + warp;
    if (w >= W) return;
    double like = 0.0;
    for (int b = 0; b < d.nb; ++b) {
        double t1 = 0.0, t2 = 0.0;
        for (int s = lane; s < d.na; s += 32) {
            double c = __ldg(d.cts + (size_t)b * d.na + s);
            double m = pred[((size_t)w * d.nb + b) * d.na + s];
            if (c == c) { t1 += c * log(m); t2 += m; }
        }
        double lb = warp_sum(t1) - warp_sum(t2);
        like += isfinite(lb) ? lb : jx_neg_inf();
    }
    if (lane == 0) cash[w] = like;
}

}  // namespace

cudaError_t jx_launch_cash(const jx_dev& d, const double* pred, int W, double* cash, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k4_cash_kernel<<<(W + K4_WARPS - 1) / K4_WARPS, K4_WARPS * 32, 0, st>>>(d, pred, W, cash);
    return cudaGetLastError();
}

cudaError_t jx_launch_xray(const jx_dev& d, const double* theta, const double* ne_ann, const double* tx_ann,
                           int W, double* pred, double* cash, uint32_t* flags, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k4_args a{d, theta, ne_ann, tx_ann, W, pred, cash, flags};
    if (d.na <= 16) {
        const size_t smem = (size_t)K4_WARPS * 2 * d.na * sizeof(double);
        const int blocks = (W + 2 * K4_WARPS - 1) / (2 * K4_WARPS);
        k4_xray_kernel<2><<<blocks, K4_WARPS * 32, smem, st>>>(a);
    } else {
        const size_t smem = (size_t)K4_WARPS * d.na * sizeof(double);
        const int blocks = (W + K4_WARPS - 1) / K4_WARPS;
        k4_xray_kernel<1><<<blocks, K4_WARPS * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}
