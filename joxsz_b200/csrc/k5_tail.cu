// K5 -- the per-walker tail of the SZ likelihood and the final sum, one warp per walker.
//
// Replaces reference joxsz_funcs.py:469-479 and :536-538 for every walker of the batch:
//
//   T0     = h(0), the not-a-knot cubic through (+-r_pp[:sep], T_SZ) at 0        (:469-471, fixed operator w_t0)
//   bright = map_out[N//2, N//2:] * convert([T0, T_SZ]) * calibration            (:472-473)
//   model  = cubic spline through (radius[sep:], bright) at the data radii       (:476, fixed operator g_op)
//   chisq  = nansum(((flux - model) / err)^2);  ll = (xray + prior) - chisq / 2  (:478-479, :536-538)
//   with calc_integ: ll -= nansum(((cint - mu) / sig)^2) / 2, cint from K1       (:480-485)
//
// `row` = map_out[N//2, N//2:] comes from the filter GEMM (k7_filter.cu) as `nparts` K-split partial rows that
// are added here in order, or from the large-map kernel's G vector through one small GEMM (jx_api.cu).
// Walkers whose status bits are set get ll = -inf (the reference returns before / regardless of this
// stage, joxsz_funcs.py:519-520, 523-525, 529-532, 536).
#include "jx_common.cuh"

namespace {

constexpr int K5_WARPS = 8, K5_WARPS_SMALL = 2, K5_SMALL_BATCH = 16384;     // walkers per CTA: fewer for small batches (even SM load)

struct k5_args {
    jx_dev d;
    const double *theta, *row, *tsz, *prior, *xlike, *cint;
    const uint32_t* flags;
    int W, ld_row, nparts;
    double *bright, *model, *chisq, *ll, *row_out;
};

// scipy interp1d(kind='linear', fill_value='extrapolate') on a small table
JX_D double linear_extrap(double x, const double* __restrict__ xk, const double* __restrict__ yk, int n) {
    int idx = 0;
    for (int i = 0; i < n; ++i) idx += (__ldg(xk + i) < x) ? 1 : 0;    // searchsorted(side='left')
    if (x != x) idx = n;
    idx = idx < 1 ? 1 : (idx > n - 1 ? n - 1 : idx);
    double x0 = __ldg(xk + idx - 1), x1 = __ldg(xk + idx), y0 = __ldg(yk + idx - 1), y1 = __ldg(yk + idx);
    double slope = (y1 - y0) / (x1 - x0);
    return slope * (x - x0) + y0;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) k5_tail_kernel(const __grid_constant__ k5_args a) {
    extern __shared__ double k5_smem[];
    const jx_dev& d = a.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * NW + warp;
    if (w >= a.W) return;                       // whole warps leave; no block barrier below
    const int H = d.nh;
    double* bright_s = k5_smem + (size_t)warp * H;

    if (a.flags && a.flags[w] != 0u) {
        if (lane == 0 && a.ll) a.ll[w] = jx_neg_inf();
        return;
    }
    const double* tsz = a.tsz + (size_t)w * d.nt;
    double s = 0.0;
    for (int i = lane; i < d.nt; i += 32) s += __ldg(d.w_t0 + i) * tsz[i];
    const double T0 = warp_sum(s);

    const int csrc = d.slot_src[JX_CALIB];
    const double calib = csrc < 0 ? d.slot_val[JX_CALIB] : a.theta[(size_t)w * d.ndim + csrc];
    const double* row = a.row + (size_t)w * a.ld_row;
    const size_t part_stride = (size_t)a.W * a.ld_row;
    for (int v = lane; v < H; v += 32) {
        const double T = v == 0 ? T0 : tsz[v - 1];
        double r = row[v];
        for (int p = 1; p < a.nparts; ++p) r += row[p * part_stride + v];
        if (a.row_out) a.row_out[(size_t)w * H + v] = r;
        const double br = r * linear_extrap(T, d.conv_T, d.conv_I, d.nconv) * calib;
        bright_s[v] = br;
        if (a.bright) a.bright[(size_t)w * H + v] = br;
    }
    __syncwarp();

    // lane = data point: model_d = sum_v g_op[d, v] bright[v]   (g_op transposed on upload: coalesced over d)
    double c = 0.0;
    for (int dpt = lane; dpt < d.nd; dpt += 32) {
        double m0 = 0.0, m1 = 0.0;
        int v = 0;
        for (; v + 1 < H; v += 2) {
            m0 += __ldg(d.g_op_t + (size_t)v * d.nd + dpt) * bright_s[v];
            m1 += __ldg(d.g_op_t + (size_t)(v + 1) * d.nd + dpt) * bright_s[v + 1];
        }
        if (v < H) m0 += __ldg(d.g_op_t + (size_t)v * d.nd + dpt) * bright_s[v];
        const double m = m0 + m1;
        if (a.model) a.model[(size_t)w * d.nd + dpt] = m;
        double z = (__ldg(d.flux + dpt) - m) / __ldg(d.flux_err + dpt);
        z = z * z;
        if (z == z) c += z;                     // np.nansum drops NaN terms
    }
    c = warp_sum(c);
    if (lane == 0) {
        if (a.chisq) a.chisq[w] = c;
        if (a.ll) {
            const double xl = a.xlike ? a.xlike[w] : 0.0;
            const double pr = a.prior ? a.prior[w] : 0.0;
            double sz_ll = -c / 2.0;
            if (d.calc_integ && a.cint) {       // joxsz_funcs.py:484-485; nansum drops a NaN term
                double z = (a.cint[w] - d.integ_mu) / d.integ_sig;
                z = z * z;
                if (z == z) sz_ll -= z / 2.0;
            }
            a.ll[w] = (xl + pr) + sz_ll;
        }
    }
}

}  // namespace

cudaError_t jx_launch_tail(const jx_dev& d, const double* theta, const double* row, int ld_row, int nparts,
                           const double* tsz, const uint32_t* flags, const double* prior, const double* xlike,
                           const double* cint, int W, double* bright, double* model, double* chisq, double* ll,
                           double* row_out, cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k5_args a{d, theta, row, tsz, prior, xlike, cint, flags, W, ld_row, nparts, bright, model, chisq, ll, row_out};
    if (W <= K5_SMALL_BATCH) {
        constexpr int NW = K5_WARPS_SMALL;
        k5_tail_kernel<NW><<<(W + NW - 1) / NW, NW * 32, (size_t)NW * d.nh * sizeof(double), st>>>(a);
    } else {
        constexpr int NW = K5_WARPS;
        k5_tail_kernel<NW><<<(W + NW - 1) / NW, NW * 32, (size_t)NW * d.nh * sizeof(double), st>>>(a);
    }
    return cudaGetLastError();
}
