// K1 -- radial profiles, priors and the hydrostatic-mass veto, one warp per walker.
//
// Replaces, per walker: Fit.updateThawed (joxsz_funcs.py:516), the parameter priors (:518-520), the
// mass veto (:522-525 with mass_fun :428-437), press_fun on r_pp (:453), temp_fun on r_pp[:sep] (:469)
// and ModelNullPot.computeProfs at the annulus mid-points (:527, :338-339).
//
// Mapping: CTA = 8 warps = 8 walkers.  Lanes stride the radial grid, so pp/T_SZ stores are coalesced
// 256-byte rows; the grid r_pp is read through the read-only path (shared by every warp).  The mass
// profile of the walker stays in shared memory for the finite-difference sign test.
#include "jx_physics.cuh"

namespace {

// warps (= walkers) per CTA: 8 for large batches; 4 for the small per-rank batches of a multi-GPU step, where 8-walker CTAs
// leave the SMs unevenly loaded (4 096 walkers = 512 CTAs on 148 SMs: 3 or 4 per SM)
constexpr int K1_WARPS = 8, K1_WARPS_SMALL = 4, K1_SMALL_BATCH = 16384;

struct k1_args {
    jx_dev d;
    const double* theta;
    int W, ld_pp;
    double *pp, *tsz, *ne_ann, *tx_ann, *prior, *cint;
    uint32_t* flags;
};

template <int NW>
__global__ void __launch_bounds__(NW * 32, 32 / NW) k1_profiles_kernel(const __grid_constant__ k1_args a) {
    extern __shared__ double k1_smem[];
    const jx_dev& d = a.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * NW + warp;
    double* mass_s = k1_smem + (size_t)warp * (d.nr + JX_NPAR + 1);
    double* par_s = mass_s + d.nr;
    if (w >= a.W) return;   // whole warp leaves together; no block-level barrier below

    // ---- scatter theta into the full parameter vector; priors on the theta columns
    double th = 0.0, pr = 0.0;
    bool bad = false;
    if (lane < d.ndim) {
        th = a.theta[(size_t)w * d.ndim + lane];
        double pa = d.prior_a[lane], pb = d.prior_b[lane];
        if (d.prior_kind[lane] == 0) {
            bad = (th < pa) || (th > pb);
        } else if (pb > 0.0) {
            double z = (th - pa) / pb;
            pr = -0.5 * log(2.0 * M_PI) - log(pb) - 0.5 * z * z;
        }
    }
    {
        const int src = lane < JX_NPAR ? d.slot_src[lane] : -1;
        const double tv = __shfl_sync(0xffffffffu, th, src < 0 ? 0 : src);
        if (lane < JX_NPAR) par_s[lane] = src < 0 ? d.slot_val[lane] : tv;
    }
    double prior_sum = warp_sum(pr) + d.prior_const;
    uint32_t flags = 0;
    if (__any_sync(0xffffffffu, bad) || !isfinite(prior_sum)) flags |= JX_FLAG_PRIOR;
    __syncwarp();

    // the radius-independent quantities of the walker are the same for every lane: one copy per warp in shared memory
    // (broadcast loads) instead of 26 doubles of registers per thread, which buys a third CTA per SM
    __shared__ jx_walker_pars wp_all[NW];
    if (lane == 0) wp_all[warp] = jx_prepare(par_s, d.dens_mode);
    __syncwarp();
    const jx_walker_pars& wp = wp_all[warp];
    if (wp.rc > wp.rs) flags |= JX_FLAG_RCRS;

    // ---- radial grid: pressure, T_SZ, mass
    double ci = 0.0;        // integrated Compton parameter: a fixed linear functional of the pressure profile
    for (int i = lane; i < d.nr; i += 32) {
        const double r = __ldg(d.r_pp + i), lr = __ldg(d.ln_r_pp + i);
        double p, dp;
        jx_pressure(wp, r, lr, p, dp);
        ci += __ldg(d.w_integ + i) * p;
        const double ne = jx_density(wp, r, lr);
        if (a.pp) a.pp[(size_t)w * a.ld_pp + i] = p;
        if (a.tsz && i < d.nt) a.tsz[(size_t)w * d.nt + i] = p / ne;
        mass_s[i] = jx_mass(dp, ne, r, 0.61);
    }
    if (a.pp)   // zero the K padding the projection GEMM reads
        for (int i = d.nr + lane; i < a.ld_pp; i += 32) a.pp[(size_t)w * a.ld_pp + i] = 0.0;
    __syncwarp();
    if (d.exclude_mass) {
        // all(np.gradient(m, 1) > 0): one-sided at the ends, central inside
        bool ok = true;
        for (int i = lane; i < d.nr; i += 32) {
            double g;
            if (i == 0) g = (mass_s[1] - mass_s[0]) / 1.0;
            else if (i == d.nr - 1) g = (mass_s[i] - mass_s[i - 1]) / 1.0;
            else g = (mass_s[i + 1] - mass_s[i - 1]) / 2.0;
            ok = ok && (g > 0.0);
        }
        if (!__all_sync(0xffffffffu, ok)) flags |= JX_FLAG_MASS;
    }

    // ---- annulus mid-points: n_e and T_X
    for (int i = lane; i < d.na; i += 32) {
        const double lr = __ldg(d.ln_midpt + i);
        const double ne = jx_density(wp, __ldg(d.midpt_kpc + i), lr);
        const double tsz = jx_pressure_only(wp, lr) / ne;
        if (a.ne_ann) a.ne_ann[(size_t)w * d.na + i] = ne;
        if (a.tx_ann) a.tx_ann[(size_t)w * d.na + i] = tsz * wp.tratio;
    }
    ci = warp_sum(ci);
    if (lane == 0) {
        if (a.flags) a.flags[w] = flags;
        if (a.prior) a.prior[w] = prior_sum;
        if (a.cint) a.cint[w] = ci;
    }
}

// ---- component methods at arbitrary radii (jx_radial_profiles)
struct kr_args {
    const double* pars;
    int W, dens_mode, n;
    const double* r;
    int r_per_walker;
    double mu_gas;
    double *press, *dpress, *ne, *tsz, *tx, *mass;
};

__global__ void __launch_bounds__(256) k_radial_kernel(const __grid_constant__ kr_args a) {
    const int w = blockIdx.x;
    __shared__ jx_walker_pars wp_s;
    if (threadIdx.x == 0) wp_s = jx_prepare(a.pars + (size_t)w * JX_NPAR, a.dens_mode);
    __syncthreads();
    const jx_walker_pars wp = wp_s;
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        const double r = a.r[(a.r_per_walker ? (size_t)w * a.n : 0) + i], lr = log(r);
        size_t o = (size_t)w * a.n + i;
        double p, dp;
        jx_pressure(wp, r, lr, p, dp);
        if (a.press) a.press[o] = p;
        if (a.dpress) a.dpress[o] = dp;
        if (a.ne || a.tsz || a.tx || a.mass) {
            double ne = jx_density(wp, r, lr);
            if (a.ne) a.ne[o] = ne;
            if (a.tsz) a.tsz[o] = p / ne;
            if (a.tx) a.tx[o] = (p / ne) * wp.tratio;
            if (a.mass) a.mass[o] = jx_mass(dp, ne, r, a.mu_gas);
        }
    }
}

}  // namespace

// per device, at jx_create: opt in to more than 48 KB of dynamic shared memory when the radial grid needs it
cudaError_t jx_profiles_configure(const jx_dev& d) {
    const size_t smem = (size_t)K1_WARPS * (d.nr + JX_NPAR + 1) * sizeof(double);
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k1_profiles_kernel<K1_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k1_profiles_kernel<K1_WARPS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return e;
}

cudaError_t jx_launch_profiles(const jx_dev& d, const double* theta, int W, double* pp, int ld_pp, double* tsz,
                               double* ne_ann, double* tx_ann, uint32_t* flags, double* prior, double* cint,
                               cudaStream_t st) {
    if (W <= 0) return cudaSuccess;
    k1_args a{d, theta, W, ld_pp, pp, tsz, ne_ann, tx_ann, prior, cint, flags};
    if (W <= K1_SMALL_BATCH) {
        constexpr int NW = K1_WARPS_SMALL;
        const size_t smem = (size_t)NW * (d.nr + JX_NPAR + 1) * sizeof(double);
        k1_profiles_kernel<NW><<<(W + NW - 1) / NW, NW * 32, smem, st>>>(a);
    } else {
        constexpr int NW = K1_WARPS;
        const size_t smem = (size_t)NW * (d.nr + JX_NPAR + 1) * sizeof(double);
        k1_profiles_kernel<NW><<<(W + NW - 1) / NW, NW * 32, smem, st>>>(a);
    }
    return cudaGetLastError();
}

extern "C" int jx_radial_profiles(const double* pars, int32_t W, int32_t dens_mode, const double* r, int32_t n,
                                  int32_t r_per_walker, double mu_gas, double* press, double* dpress, double* ne, double* tsz,
                                  double* tx, double* mass, int32_t device, void* stream) {
    if (!pars || !r || W <= 0 || n <= 0) return JX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    kr_args a{pars, W, dens_mode, n, r, r_per_walker, mu_gas, press, dpress, ne, tsz, tx, mass};
    k_radial_kernel<<<W, 256, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}
