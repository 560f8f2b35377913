// Internal declarations shared by the kernels of libjoxsz_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>

#include "../../include/joxsz_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libjoxsz_b200 is written for sm_100a (B200) only"
#endif

#define JX_HD __host__ __device__ __forceinline__
#define JX_D __device__ __forceinline__

constexpr int JX_WARP = 32;
constexpr int JX_MAX_NDIM = 32;   // theta columns handled by one warp
constexpr int JX_BMIX_ROWS = 28, JX_BMIX_PITCH = 132;   // device layout of jx_dev::bmix (K3 direct phase B)

// One quarter-plane pixel of the Compton-y map, for the synthesis phase of K3: the spline piece that
// covers its distance from the centre and the offset from that piece's left knot.  (u, v) and (v, u)
// share the value (the map is radial), so only u <= v is stored.
struct __align__(16) jx_synth_px {
    double dx;
    uint16_t seg, u, v, pad;
};

// Device-resident constants + workspace of one handle.
struct jx_dev {
    // parameters
    int ndim, dens_mode, exclude_mass;
    int slot_src[JX_NPAR];
    double slot_val[JX_NPAR];
    const int32_t* prior_kind;
    const double *prior_a, *prior_b;
    double prior_const;
    // SZ geometry
    int nr, nrp /* nr rounded up to 8: leading dimension of ws_pp and the operators */, nt, nmap, nh, npad, nq, nseg,
        ncoef /* 4*nseg */;
    const double* r_pp;
    const double *ln_r_pp, *ln_midpt;   // natural logs of r_pp / midpt_kpc (derived at jx_create)
    const double* proj_op;   // [ncoef, nrp] zero padded, rows in two planes [c0 c1 per piece][c2 c3 per piece] (production layout)
    const double* proj_op_tap; // [ncoef, nrp] rows [4][seg] as supplied (jx_sz_project's `coef` output)
    const double* y_op;      // [nr, nrp] zero padded
    const int32_t* seg;      // [nh, nh]
    const double* dx;        // [nh, nh]
    const double* bhat;      // [nq, nq]
    int nbeam;               // beam half-side incl. the centre
    int k3_direct;           // map kernel convolves along y directly (jx_szmap_direct_ok at jx_create)
    int k3_ws;               // two walkers in flight per SM: the warp-specialised map kernel (k3w_szmap.cu)
    int k3l2;                // large maps: the two-CTA-per-SM form of the L2-staged kernel (k3l2_szmap.cu)
    const double* bmix;      // [28, bmix_pitch] beam in (y offset, kx), zero rows beyond nbeam; NULL when nbeam > 28
    int bmix_pitch;          // JX_BMIX_PITCH for the cyclic length 256, nq rounded up to 4 otherwise
    const double* cmat_t;    // [nh, nh] transposed on upload: [kx, v]
    const double* hf;        // [nh, nh] [u, kx]
    const double* dinv;      // [nh, nh] [kx, v]
    const double* filt_q;    // [nh, nh]
    // derived at jx_create
    int hp8, hp16;           // nh rounded up to 8 / 16
    int xs_pitch;            // large-map path: doubles per row of the per-CTA scratch map (multiple of 4, >= nq and hp16)
    double* ws_scratch;      // large-map path: [sm_count][hp8][xs_pitch]
    double* ws_scratch2;     // large-map path, direct y convolution: its output map, same shape
    const uint16_t* seg16;   // [nh, nh] seg narrowed
    const double* costab;    // [nmap] cos(2 pi m / nmap)
    const jx_synth_px* synth; // [nsynth] quarter-plane pixels with u <= v, padded with u = 0xffff sentinels
    int nsynth;               // multiple of 256
    const jx_synth_px* synth_tiles;  // large maps only: the whole quarter plane in 32 x 32 tiles (ub <= vb, row-major), padded with sentinels
    int nsynth_tiles;                // number of tiles (1024 entries each)
    const double2* bhat_sw;  // [ceil(nq/2)][16][9] beam spectrum of column pair cp in FFT thread order (K3 phase B: position p, thread t = 0..8)
    const double* w_t0;      // [nt]
    int nconv;
    const double *conv_T, *conv_I;
    int nd;
    const double *g_op, *flux, *flux_err;
    int calc_integ;
    const double* w_integ;   // [nr] (zeros when the operator was not supplied)
    double integ_mu, integ_sig;
    const double* g_op_t;    // [nh, nd] g_op transposed (K5 reads it coalesced over the data points)
    // filter stage as one GEMM over the walkers (cyclic length 256; k7_filter.cu)
    int hpf;                 // leading dimension of filt_op's output index and of the partial rows (>= hp8)
    int ntri, ktri;          // nh (nh + 1) / 2 pixels u <= v of the quarter plane; rounded up to 32
    const double* filt_op;   // [hpf, ktri] zero padded: filt_op[x, (u,v)] = response of map_out[N//2, N//2 + x] to conv_c[u,v]
    // X-ray
    int na, nb, ntab;
    const double *midpt_kpc, *projvols, *tlog, *lnrate0, *lnrate1, *cts, *srcscale, *bkgterm;
    double tmin, tmax;
    // workspace [max_walkers, ...]
    int max_walkers;
    double* ws_pp;      // [W, nrp] zero padded
    double* ws_tsz;     // [W, nt]
    double* ws_ne;      // [W, na]
    double* ws_tx;      // [W, na]
    double* ws_prior;   // [W]
    double* ws_integ;   // [W] integrated Compton parameter (K1)
    double* ws_xlike;   // [W]
    uint32_t* ws_flags; // [W]
    double* ws_coef;    // [W, ncoef]
    double* ws_tri;     // [W, ktri] packed triangle of the convolved map (map kernel -> filter GEMM), zero padded
    double* ws_rowp;    // [jx_filter_parts(d) * W, hpf] K-split partial sums of the filter GEMM
    double* ws_convq;   // tap only, allocated lazily: [W, nh, nh]
};

struct jx_handle {
    jx_dev d;
    int device;
    int sm_count;
    std::string err;
    void* allocs[64];
    int nallocs;
    size_t convq_capacity;   // walkers for which ws_convq is allocated
    double* tap_scratch;     // full-map tap scratch [W, 2, nh, nh]
    size_t tap_scratch_walkers;
    // profiling
    int profiling;
    cudaEvent_t ev[JX_NSTAGE + 1];   // ev[0..6] on the caller's stream; the X-ray stage is bracketed by evx[0..1] on the side stream
    cudaEvent_t evx[2];
    bool ev_ready;
    double stage_ms[JX_NSTAGE];
    int64_t stage_launches[JX_NSTAGE];
    bool pending;            // events recorded but not yet accumulated
    // the X-ray kernel only depends on the profiles: it runs on a side stream next to the projection GEMM
    cudaStream_t side;
    cudaEvent_t ev_fork, ev_join;
    int pending_launches[JX_NSTAGE];
    double* collapsed_op;    // [hp8, nrp]: row = collapsed_op . pp (collapsed mode, built on first use)
};

// ---- launchers implemented by the kernel files (all asynchronous on `st`)
cudaError_t jx_launch_profiles(const jx_dev& d, const double* theta, int W, double* pp, int ld_pp, double* tsz,
                               double* ne_ann, double* tx_ann, uint32_t* flags, double* prior, double* cint,
                               cudaStream_t st);
cudaError_t jx_launch_project(const jx_dev& d, const double* pp, int W, const double* op, int nout,
                              double* out, cudaStream_t st);
cudaError_t jx_launch_xray(const jx_dev& d, const double* theta, const double* ne_ann, const double* tx_ann,
                           int W, double* pred, double* cash, uint32_t* flags, cudaStream_t st);
cudaError_t jx_launch_cash(const jx_dev& d, const double* pred, int W, double* cash, cudaStream_t st);
cudaError_t jx_profiles_configure(const jx_dev& d);   // one-time per device (jx_create)
cudaError_t jx_gemm_configure();
cudaError_t jx_launch_gemm_nt(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N,
                              int Kpad, cudaStream_t st);
// production map stage: coef -> G[kx] (+ optional quarter-plane convolved map).  `flags` may be NULL
// (evaluate every walker); flagged walkers are skipped.
cudaError_t jx_launch_szmap(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                            double* convq, double* tri, cudaStream_t st);
// filter stage (k7_filter.cu): rowp[kparts][W][hp8] = K-split partial sums of tri[W, ktri] . filt_op^T
constexpr int JX_FILTER_CPP = 13;         // chunks of 32 packed pixels per K part
int jx_filter_pitch(int hp8);
cudaError_t jx_filter_configure(const jx_dev& d);
int jx_filter_parts(const jx_dev& d);     // ws_rowp holds jx_filter_parts(d) * max_walkers rows
cudaError_t jx_launch_filter(const jx_dev& d, const double* tri, int W, double* rowp, cudaStream_t st);
// tail: row -> bright, model, chisq, ll (any output may be NULL).  row = sum of `nparts` partial rows,
// part p at row + p * W * ld_row.
cudaError_t jx_launch_tail(const jx_dev& d, const double* theta, const double* row, int ld_row, int nparts,
                           const double* tsz, const uint32_t* flags, const double* prior, const double* xlike,
                           const double* cint, int W, double* bright, double* model, double* chisq, double* ll,
                           double* row_out, cudaStream_t st);
bool jx_szmap_direct_ok(const jx_dev& d);
cudaError_t jx_szmap_configure(const jx_dev& d);   // one-time cudaFuncSetAttribute
cudaError_t jx_launch_tap_y2d(const jx_dev& d, const double* coef, int W, double* y2d, cudaStream_t st);
cudaError_t jx_launch_tap_expand(const jx_dev& d, const double* convq, int W, double* conv2d, cudaStream_t st);
cudaError_t jx_launch_tap_mapout(const jx_dev& d, const double* convq, int W, double* mapout, double* scratch,
                                 cudaStream_t st);
size_t jx_szmap_smem_bytes(const jx_dev& d);
// warp-specialised form of the map stage, two walkers in flight per SM (k3w_szmap.cu)
bool jx_szmap_ws_ok(const jx_dev& d);
size_t jx_szmap_ws_smem_bytes(const jx_dev& d);
cudaError_t jx_szmap_ws_configure(const jx_dev& d);
cudaError_t jx_launch_szmap_ws(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                               double* convq, double* tri, cudaStream_t st);
// large-map path (cyclic length 512 / 1024, working set in an L2-resident global scratch): k3l_szmap.cu
bool jx_szmap_large_supported(const jx_dev& d);
size_t jx_szmap_large_smem_bytes(const jx_dev& d);
cudaError_t jx_szmap_large_configure(const jx_dev& d);
cudaError_t jx_launch_szmap_large(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                                  double* convq, double* tri, double* scratch, double* scratch2, cudaStream_t st);

// two CTAs per SM, transforms straight from / to the L2 scratch maps (k3l2_szmap.cu)
bool jx_szmap_large2_ok(const jx_dev& d);
size_t jx_szmap_large2_smem_bytes(const jx_dev& d);
cudaError_t jx_szmap_large2_configure(const jx_dev& d);
cudaError_t jx_launch_szmap_large2(const jx_dev& d, const double* coef, const uint32_t* flags, int W, int sm_count,
                                   double* convq, double* tri, double* scratch, double* scratch2, cudaStream_t st);

// ---- small device helpers
JX_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
JX_D double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
JX_D double jx_neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }
