/* joxsz_b200.h -- C ABI of libjoxsz_b200.so: the batched JoXSZ joint SZ + X-ray log-likelihood
 * on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of fcastagna/JoXSZ: the per-walker likelihood that
 * emcee calls once per walker per step (reference joxsz_funcs.py:507-546 `getLikelihood`, bound onto
 * mbproj2's Fit at joxsz_main.py:186-188 and handed to emcee at joxsz_main.py:206).  The reference has
 * no FFI of its own (it is pure Python); the Python binding a maintainer adds is a ctypes stub, shown
 * in INTEGRATION.md.  Everything below takes plain pointers and sizes; no torch / CUDA types appear.
 *
 * Conventions
 *   - every `double*` / `int32_t*` argument of a compute call is a DEVICE pointer on the handle's
 *     device unless marked [host]; arrays are C order (row-major), float64;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous on it;
 *   - every function returns 0 on success or a negative jx_status; nothing throws; the message of
 *     the last failure is available from jx_last_error();
 *   - a handle is not thread-safe; one handle per (device, stream);
 *   - the library never allocates per call: workspace is sized at jx_create for `max_walkers`;
 *   - there is no CPU fallback: jx_create fails if no sm_100-class device is present.
 */
#ifndef JOXSZ_B200_H
#define JOXSZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JX_ABI_VERSION 7

typedef enum jx_status {
    JX_OK = 0,
    JX_ERR_INVALID = -1,   /* bad argument / unsupported geometry */
    JX_ERR_CUDA = -2,      /* CUDA runtime error (see jx_last_error) */
    JX_ERR_NO_DEVICE = -3, /* no usable sm_100 device */
    JX_ERR_CAPACITY = -4   /* W exceeds max_walkers */
} jx_status;

/* Slots of the full physical parameter vector (names are the keys of the reference's `fit.pars`;
 * defaults/bounds: joxsz_funcs.py:266-272, 318, 358-372; joxsz_main.py:131,156-157). */
enum jx_param_slot {
    JX_P0 = 0,      /* "P_0"              gNFW normalisation, keV cm^-3 */
    JX_A,           /* "a"                                                 */
    JX_B,           /* "b"                                                 */
    JX_C,           /* "c"                                                 */
    JX_RP,          /* "r_p"              kpc                              */
    JX_LOGN0,       /* "log(n_0)"         log10 cm^-3                      */
    JX_BETA,        /* "\\beta"                                            */
    JX_LOGRC,       /* "log(r_c)"         log10 kpc                        */
    JX_LOGRS,       /* "log(r_s)"         log10 kpc                        */
    JX_ALPHA,       /* "\\alpha"                                           */
    JX_EPS,         /* "\\epsilon"                                         */
    JX_GAMMA,       /* "\\gamma"                                           */
    JX_LOGN02,      /* "log(n_{02})"      (density mode 'double' only)     */
    JX_BETA2,       /* "\\beta_2"                                          */
    JX_LOGRC2,      /* "log(r_{c2})"                                       */
    JX_LOGTRATIO,   /* "log(T_X/T_{SZ})"                                   */
    JX_ZMET,        /* "Z"                solar                            */
    JX_BACKSCALE,   /* "backscale"                                         */
    JX_CALIB,       /* "calibration"                                       */
    JX_NPAR
};

/* Per-walker status bits (jx_profiles `flags`): any bit set => log-likelihood is -inf. */
#define JX_FLAG_PRIOR   1u /* a parameter prior is -inf or not finite   (joxsz_funcs.py:518-520) */
#define JX_FLAG_MASS    2u /* hydrostatic mass not monotone             (joxsz_funcs.py:522-525) */
#define JX_FLAG_RCRS    4u /* r_c > r_s                                 (joxsz_funcs.py:397-407, 536) */
#define JX_FLAG_XNONPOS 8u /* an X-ray predicted profile is not > 0     (joxsz_funcs.py:529-532) */

/* Everything jx_create needs, as [host] pointers; the library copies what it keeps.
 * The Python side fills this from a built `fit` object (joxsz_b200/packer.py). */
typedef struct jx_setup {
    int32_t abi_version;             /* must be JX_ABI_VERSION */
    int32_t device;                  /* CUDA device ordinal */
    int32_t max_walkers;             /* workspace capacity */

    /* -- parameters: theta[W, ndim] -> full vector (replaces Fit.updateThawed, joxsz_funcs.py:516) */
    int32_t ndim;                    /* number of thawed (sampled) parameters */
    int32_t slot_src[JX_NPAR];       /* >= 0: column of theta feeding this slot; -1: frozen */
    double  slot_val[JX_NPAR];       /* value used when frozen */
    int32_t dens_mode;               /* 0 = 'single', 1 = 'double' (joxsz_funcs.py:390-394) */
    int32_t exclude_unphy_mass;      /* joxsz_main.py:88 */
    /* priors, one per theta column (mbproj2 Param / ParamGaussian .prior()) */
    const int32_t* prior_kind;       /* [ndim] 0 = box, 1 = Gaussian */
    const double*  prior_a;          /* [ndim] minval | mu */
    const double*  prior_b;          /* [ndim] maxval | sigma */
    double prior_const;              /* sum of the (finite) priors of frozen parameters */

    /* -- SZ geometry (SZ_data, joxsz_funcs.py:136-170) */
    int32_t nr;                      /* len(r_pp) */
    int32_t nt;                      /* sep: T_SZ is evaluated on r_pp[:nt] (joxsz_funcs.py:469) */
    int32_t nmap;                    /* N: side of d_mat / filtering (odd) */
    int32_t nh;                      /* H = N/2 + 1 */
    int32_t npad;                    /* P: cyclic length of the beam convolution */
    int32_t nseg;                    /* spline pieces referenced by the map */
    const double*  r_pp;             /* [nr] kpc */
    const double*  proj_op;          /* [4*nseg, nr]  pressure -> spline-piece coefficients of the Compton-y profile
                                        (Abel transform * y scaling * not-a-knot fit; joxsz_funcs.py:457-460) */
    const double*  y_op;             /* [nr, nr]      pressure -> Compton-y profile (joxsz_funcs.py:457-459) */
    const int32_t* seg;              /* [H, H] spline piece of quarter-plane pixel (u, v) */
    const double*  dx;               /* [H, H] offset from the piece's left knot */
    const double*  bhat;             /* [P/2+1, P/2+1] beam spectrum * step^2 / P^2 */
    const double*  cmat;             /* [H, H]  [v, kx]  w_v cos(2 pi kx v / N) */
    const double*  hf;               /* [H, H]  [u, kx]  w_u sum_ky filt[ky,kx] cos(2 pi ky u / N) */
    const double*  dinv;             /* [H, H]  [kx, v]  w_kx cos(2 pi kx v / N) / N^2 */
    const double*  filt_q;           /* [H, H]  filtering[:H, :H] (full-map tap only) */
    int32_t nbeam;                   /* B/2 + 1: half-side of the beam image incl. the centre (joxsz_funcs.py:55-77) */
    const double*  bmix;             /* [nbeam, P/2+1] beam in (y offset j, x frequency kx) * step^2 / P: the map
                                        kernel convolves along y directly when nbeam <= 28 and H <= 88 */
    /* -- SZ tail (joxsz_funcs.py:469-479) */
    const double*  w_t0;             /* [nt]  h(0) = w_t0 . t_prof */
    int32_t nconv;
    const double*  conv_T;           /* [nconv] keV   (joxsz_main.py:108-109) */
    const double*  conv_I;           /* [nconv] 1e3 * I0 */
    int32_t nd;                      /* SZ data points */
    const double*  g_op;             /* [nd, H] spline through (radius[sep:], prof) evaluated at the data radii */
    const double*  flux;             /* [nd] */
    const double*  flux_err;         /* [nd] */
    /* -- optional integrated-Compton-parameter penalty (joxsz_funcs.py:480-487, `calc_integ`, joxsz_main.py:65) */
    int32_t calc_integ;              /* 0 = off (reference default) */
    const double*  w_integ;          /* [nr] cint = w_integ . pressure  (Simpson weights * 2 pi * y scaling * Abel), may be
                                        NULL when calc_integ == 0 */
    double integ_mu, integ_sig;      /* Gaussian penalty -((cint - mu) / sig)^2 / 2 */

    /* -- X-ray (mbproj2 Annuli / Band / CountRate; joxsz_funcs.py:184-211, 495-505) */
    int32_t na;                      /* annuli = shells */
    int32_t nb;                      /* energy bands */
    int32_t ntab;                    /* temperature-table length */
    const double*  midpt_kpc;        /* [na] */
    const double*  projvols;         /* [na, na] annulus x shell, cm^3 */
    const double*  tlog;             /* [ntab] ln T grid */
    double tmin, tmax;               /* clip of T before ln */
    const double*  lnrate0;          /* [nb, ntab] ln rate at Z = 0 */
    const double*  lnrate1;          /* [nb, ntab] ln rate at Z = 1 */
    const double*  cts;              /* [nb, na] observed counts; NaN = missing bin */
    const double*  srcscale;         /* [nb, na] areascale * exposure */
    const double*  bkgterm;          /* [nb, na] backrate * geomarea * areascale * exposure (x backscale) */
} jx_setup;

typedef struct jx_handle jx_handle;

/* Build a handle: validates the geometry, copies every constant to the device, sizes the workspace. */
int  jx_create(const jx_setup* setup, jx_handle** out);
void jx_destroy(jx_handle* h);
/* Message of the last failure on this handle (h may be NULL: last jx_create failure). */
const char* jx_last_error(const jx_handle* h);

/* THE hot path: replaces one emcee batch of `getLikelihood` calls (joxsz_funcs.py:507-546).
 * theta [W, ndim] -> ll [W]; -inf exactly where the reference returns -inf; never NaN. */
int jx_loglike(jx_handle* h, const double* theta, int32_t W, double* ll, void* stream);

/* Collapsed mode (optional, `ll` only): every step between the pressure profile and the consumed row of the filtered
 * map is linear (joxsz_funcs.py:457-467), so row = L pp with a constant L [nh, nr].  L is built once per handle by
 * pushing the nr unit profiles through the staged kernels above; the call is then K1 -> one GEMM -> tail.  Same
 * inputs / outputs / -inf conventions as jx_loglike; the intermediate maps do not exist in this mode. */
int jx_loglike_collapsed(jx_handle* h, const double* theta, int32_t W, double* ll, void* stream);

/* ---- parity taps: every output pointer may be NULL (not produced).  Taps evaluate every walker,
 *      like the reference's component methods, regardless of prior flags. */

/* K1: pressure on r_pp (`get_sz_like('pp')`, joxsz_funcs.py:453), T_SZ on r_pp[:nt] (:469),
 * n_e and T_X at the annulus mid-points (:338-339, mbproj2 computeProfs), status bits, summed prior. */
int jx_profiles(jx_handle* h, const double* theta, int32_t W,
                double* pp /*[W,nr]*/, double* tsz /*[W,nt]*/, double* ne_ann /*[W,na]*/,
                double* tx_ann /*[W,na]*/, uint32_t* flags /*[W]*/, double* prior /*[W]*/, void* stream);

/* K2: Compton-y profile (joxsz_funcs.py:457-459) and the spline-piece coefficients the map uses. */
int jx_sz_project(jx_handle* h, const double* theta, int32_t W,
                  double* y /*[W,nr]*/, double* coef /*[W,4*nseg]*/, void* stream);

/* K3 full maps: y_2d (:462), conv_2d (:464), map_out (:466-467), each [W, N, N]. */
int jx_sz_maps(jx_handle* h, const double* theta, int32_t W,
               double* y2d, double* conv2d, double* mapout, void* stream);

/* K3+K5: filtered row map_out[N//2, N//2:] [W,H]; `get_sz_like('bright')` [W,H] (:472-473);
 * model at the data radii [W,nd] (:476); chisq [W] (:478); integrated Compton parameter cint [W] (:481-483,
 * `get_sz_like('integ')`; computed whether or not calc_integ is set). */
int jx_sz_profile(jx_handle* h, const double* theta, int32_t W,
                  double* row, double* bright, double* model, double* chisq, double* cint, void* stream);

/* K4: predicted X-ray profiles (`Fit.calcProfiles`, :527) [W,nb,na] and the Cash log-likelihood
 * (`mylikeFromProfs`, :495-505; -inf where a profile is not > 0, :529-532) [W]. */
int jx_xray(jx_handle* h, const double* theta, int32_t W, double* pred, double* cash, void* stream);

/* `mylikeFromProfs` on caller-supplied predicted profiles (joxsz_funcs.py:495-505): pred [W,nb,na] ->
 * cash [W] = sum over bands of cashLogLikelihood over bins whose counts are not NaN (-inf if a band's
 * sum is not finite).  No positivity gate: that is the caller's `if` at :529-532. */
int jx_cash_from_profiles(jx_handle* h, const double* pred, int32_t W, double* cash, void* stream);

/* Component methods at arbitrary radii (press_fun :275, press_derivative :289, vikhFunction :375,
 * temp_fun :321, mass_fun :428).  No handle: `pars` [W, JX_NPAR] holds full parameter vectors.
 * `r` is one grid [n] shared by all walkers, or [W, n] when r_per_walker != 0 (the r_500 root search of
 * joxsz_plots.py:316-339 evaluates each sample at its own radius).  Outputs [W, n] each, NULL to skip. */
int jx_radial_profiles(const double* pars, int32_t W, int32_t dens_mode, const double* r, int32_t n,
                       int32_t r_per_walker, double mu_gas, double* press, double* dpress, double* ne, double* tsz, double* tx,
                       double* mass, int32_t device, void* stream);

/* ---- ensemble stretch move (emcee RedBlueMove/StretchMove semantics; joxsz_main.py:206-210).
 * Every rank holds the whole ensemble `coords` [nall, ndim], `lp` [nall] (kept identical by the
 * all-gather of each half-step's results).  `perm` [nall] is a random permutation of 0..nall-1 shared
 * by all ranks (jx_stretch_permutation); the colour of the walker at position p of `perm` is p & 1 (emcee:
 * `inds = arange(n) % 2; shuffle(inds)`).  In the half-step `split` the active walkers are
 * perm[2 r + split], r = 0..ns-1 with ns = (nall - split + 1) / 2; a rank processes the contiguous slice
 * r in [r_first, r_first + r_count) -- equal, fixed-size work per rank whatever the colouring.
 * RNG: Philox4x32-10, key = seed, counter = (walker index, iteration lo, iteration hi, purpose | split << 2) with
 * purpose 0 = proposal, 1 = acceptance, 2 = colouring keys: every draw of a walker in an iteration comes from its own
 * Philox block, and a chain does not depend on the number of ranks.
 * `iter_dev` (device pointer, may be NULL) is added to `iteration`: with the counter in device memory one sampler
 * iteration can be captured in a CUDA graph and replayed (jx_stretch_advance increments the counter on the stream).
 *
 * propose: partner j = perm[2 rint + (1 - split)], rint uniform over the other colour;
 *          z = ((a-1) u + 1)^2 / a;  prop = c_j - (c_j - x_k) z;  factor = (ndim - 1) ln z. */
int jx_stretch_propose(const double* coords, const int32_t* perm, int32_t nall, int32_t ndim, int32_t split,
                       int32_t r_first, int32_t r_count, double a, uint64_t seed, uint64_t iteration,
                       const uint64_t* iter_dev, double* prop /*[r_count,ndim]*/, double* factor /*[r_count]*/,
                       int32_t device, void* stream);
/* accept: packed[i] = (new position [ndim], new log-prob, accepted 0/1) for slice entry i, where the
 * move is accepted iff factor + lp_new - lp[k] > ln(u) (emcee RedBlueMove.propose). */
int jx_stretch_accept(const double* coords, const double* lp, const int32_t* perm, int32_t nall, int32_t ndim,
                      int32_t split, int32_t r_first, int32_t r_count, const double* prop, const double* lp_new,
                      const double* factor, uint64_t seed, uint64_t iteration, const uint64_t* iter_dev,
                      double* packed /*[r_count, ndim+2]*/, int32_t device, void* stream);
/* permutation: perm = argsort of 64 Philox bits per walker (stable), a uniformly random permutation that
 * depends only on (seed, iteration) -- identical on every rank, generated on the device.
 * Call with workspace == NULL to get the required size in *workspace_bytes. */
int jx_stretch_permutation(int32_t* perm, int32_t nall, uint64_t seed, uint64_t iteration, const uint64_t* iter_dev,
                           void* workspace, size_t* workspace_bytes, int32_t device, void* stream);
/* *iter_dev += by, in stream order (the last node of a captured sampler iteration). */
int jx_stretch_advance(uint64_t* iter_dev, uint64_t by, int32_t device, void* stream);
/* scatter: write the gathered results of all ranks, packed_all [>= ns, ndim+2] in r order, back into
 * coords / lp and add the acceptance flags to naccept [nall]. */
int jx_stretch_scatter(double* coords, double* lp, int32_t* naccept, const int32_t* perm, int32_t nall,
                       int32_t ndim, int32_t split, const double* packed_all, int32_t ns, int32_t device,
                       void* stream);
/* accept fused with the exchange over peer memory (NVLink / NVSwitch), replacing accept + all-gather: every rank
 * holds a buffer packed_all [2][world * per0][ndim + 2] and flags [2][world] (uint64, zero-initialised) that all ranks
 * of the node can address (e.g. torch symmetric memory); `peer_packed` / `peer_flags` are HOST arrays of the `world`
 * device addresses, in rank order.  The kernel writes this rank's rows (slice entry i -> row rank * per0 + i of
 * buffer `split`) into every rank's packed_all and then stores the epoch `iteration + 1` into flags[split][rank] of
 * every rank with release semantics at system scope.  `done` is a zero-initialised device counter owned by the caller.
 * scatter_p2p waits until all `world` flags of buffer `split` reach the epoch and then applies the local copy (rank
 * g's rows are r = g * per .. in half-step order).  emcee semantics are those of jx_stretch_accept / _scatter. */
int jx_stretch_accept_p2p(const double* coords, const double* lp, const int32_t* perm, int32_t nall, int32_t ndim,
                          int32_t split, int32_t r_first, int32_t r_count, const double* prop, const double* lp_new,
                          const double* factor, uint64_t seed, uint64_t iteration, const uint64_t* iter_dev,
                          const uint64_t* peer_packed, const uint64_t* peer_flags, int32_t world, int32_t rank,
                          int32_t per0, uint32_t* done, int32_t device, void* stream);
int jx_stretch_scatter_p2p(double* coords, double* lp, int32_t* naccept, const int32_t* perm, int32_t nall,
                           int32_t ndim, int32_t split, const double* packed_local, const uint64_t* flags_local,
                           int32_t ns, int32_t world, int32_t per, int32_t per0, uint64_t iteration,
                           const uint64_t* iter_dev, int32_t device, void* stream);

/* ---- measurement helpers (bench.py) */
enum jx_stage { JX_ST_PROFILES = 0, JX_ST_PROJECT, JX_ST_SZMAP, JX_ST_XRAY, JX_ST_TAIL, JX_ST_FILTER, JX_NSTAGE };
/* When on, jx_loglike brackets each stage with CUDA events on `stream`. */
int jx_set_profiling(jx_handle* h, int32_t on);
/* Sum of per-stage device milliseconds and launch counts since the last reset [host outputs]; resets. */
int jx_stage_times(jx_handle* h, double* ms /*[JX_NSTAGE]*/, int64_t* launches /*[JX_NSTAGE]*/);
/* Library build info (arch, ABI). */
const char* jx_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* JOXSZ_B200_H */
