// K6 -- affine-invariant stretch move (Goodman & Weare 2010) with emcee's red/blue split.
//
// Replaces, for the walkers owned by this rank, emcee's StretchMove.get_proposal and the accept loop
// of RedBlueMove.propose (emcee 3.x; driven by the reference at joxsz_funcs.py:593-622 through
// EnsembleSampler, joxsz_main.py:206-210).  The likelihood of the proposals is evaluated in between
// by jx_loglike.  Random numbers come from Philox4x32-10 keyed by the run seed with the counter
// (global walker index, iteration, purpose | split << 2), so a chain is bit-identical however the
// ensemble is sharded over GPUs.  Every (purpose, split) pair has its own counter word: the shuffle keys, the
// proposal draws and the acceptance draws of a walker in one iteration are independent Philox blocks.
// The iteration every kernel uses is `iteration + *iter_dev` (iter_dev may be NULL): with the counter in device
// memory a whole sampler iteration is one CUDA graph that is replayed unchanged (jx_stretch_advance bumps it).
#include <cub/device/device_radix_sort.cuh>

#include "jx_common.cuh"

namespace {

struct philox4 { uint32_t v[4]; };

JX_HD philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

JX_HD double u01(uint32_t hi, uint32_t lo) {       // 53-bit uniform in [0, 1)
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

constexpr uint32_t PURPOSE_PROPOSE = 0u, PURPOSE_ACCEPT = 1u, PURPOSE_SHUFFLE = 2u;
JX_HD uint32_t stream_word(uint32_t purpose, int split) { return purpose | ((uint32_t)split << 2); }
JX_D uint64_t effective_iteration(uint64_t iteration, const uint64_t* iter_dev) {
    return iter_dev ? iteration + *iter_dev : iteration;
}

__global__ void k6_propose_kernel(const double* __restrict__ coords, const int32_t* __restrict__ perm, int nall,
                                  int ndim, int split, int r_first, int r_count, double a, uint64_t seed,
                                  uint64_t iteration, const uint64_t* __restrict__ iter_dev,
                                  double* __restrict__ prop, double* __restrict__ factor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r_count) return;
    iteration = effective_iteration(iteration, iter_dev);
    const int k = perm[2 * (r_first + i) + split];
    philox4 r = philox4x32_10((uint32_t)k, (uint32_t)iteration, (uint32_t)(iteration >> 32),
                              stream_word(PURPOSE_PROPOSE, split), (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = u01(r.v[0], r.v[1]);
    const double root = (a - 1.0) * u + 1.0;
    const double z = root * root / a;
    const int other = 1 - split;
    const int nc = (nall - other + 1) / 2;                // positions of parity `other` in 0..nall-1
    const int rint = (int)(((uint64_t)r.v[2] * (uint64_t)nc) >> 32);
    const double* s = coords + (size_t)k * ndim;
    const double* c = coords + (size_t)perm[2 * rint + other] * ndim;
    double* q = prop + (size_t)i * ndim;
    for (int d = 0; d < ndim; ++d) q[d] = c[d] - (c[d] - s[d]) * z;
    factor[i] = ((double)ndim - 1.0) * log(z);
}

__global__ void k6_accept_kernel(const double* __restrict__ coords, const double* __restrict__ lp,
                                 const int32_t* __restrict__ perm, int ndim, int split, int r_first, int r_count,
                                 const double* __restrict__ prop, const double* __restrict__ lp_new,
                                 const double* __restrict__ factor, uint64_t seed, uint64_t iteration,
                                 const uint64_t* __restrict__ iter_dev, double* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r_count) return;
    iteration = effective_iteration(iteration, iter_dev);
    const int k = perm[2 * (r_first + i) + split];
    philox4 r = philox4x32_10((uint32_t)k, (uint32_t)iteration, (uint32_t)(iteration >> 32),
                              stream_word(PURPOSE_ACCEPT, split), (uint32_t)seed, (uint32_t)(seed >> 32));
    const double lnu = log(u01(r.v[0], r.v[1]));
    const double lnpdiff = factor[i] + lp_new[i] - lp[k];
    const bool acc = lnpdiff > lnu;
    const double* src = acc ? prop + (size_t)i * ndim : coords + (size_t)k * ndim;
    double* o = packed + (size_t)i * (ndim + 2);
    for (int d = 0; d < ndim; ++d) o[d] = src[d];
    o[ndim] = acc ? lp_new[i] : lp[k];
    o[ndim + 1] = acc ? 1.0 : 0.0;
}

__global__ void k6_scatter_kernel(double* __restrict__ coords, double* __restrict__ lp, int32_t* __restrict__ naccept,
                                  const int32_t* __restrict__ perm, int ndim, int split,
                                  const double* __restrict__ packed_all, int ns) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ns) return;
    const int k = perm[2 * r + split];
    const double* o = packed_all + (size_t)r * (ndim + 2);
    for (int d = 0; d < ndim; ++d) coords[(size_t)k * ndim + d] = o[d];
    lp[k] = o[ndim];
    if (naccept && o[ndim + 1] != 0.0) naccept[k] += 1;
}

// sort keys of the colouring permutation: 64 random bits per walker (ties are broken by the stable sort)
__global__ void k6_shuffle_keys_kernel(uint64_t* __restrict__ keys, int32_t* __restrict__ vals, int nall, uint64_t seed,
                                       uint64_t iteration, const uint64_t* __restrict__ iter_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nall) return;
    iteration = effective_iteration(iteration, iter_dev);
    philox4 r = philox4x32_10((uint32_t)i, (uint32_t)iteration, (uint32_t)(iteration >> 32),
                              stream_word(PURPOSE_SHUFFLE, 0), (uint32_t)seed, (uint32_t)(seed >> 32));
    keys[i] = ((uint64_t)r.v[0] << 32) | r.v[1];
    vals[i] = i;
}

__global__ void k6_advance_kernel(uint64_t* iter_dev, uint64_t by) { *iter_dev += by; }

}  // namespace

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

extern "C" int jx_stretch_permutation(int32_t* perm, int32_t nall, uint64_t seed, uint64_t iteration,
                                      const uint64_t* iter_dev, void* workspace, size_t* workspace_bytes,
                                      int32_t device, void* stream) {
    if (nall < 1 || !workspace_bytes) return JX_ERR_INVALID;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, nall);
    const size_t kb = align256(sizeof(uint64_t) * (size_t)nall), vb = align256(sizeof(int32_t) * (size_t)nall);
    const size_t need = 2 * kb + vb + align256(cub_bytes);
    if (!workspace) {                 // size query
        *workspace_bytes = need;
        return JX_OK;
    }
    if (!perm || *workspace_bytes < need) return JX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)workspace;
    uint64_t* keys_in = (uint64_t*)base;
    uint64_t* keys_out = (uint64_t*)(base + kb);
    int32_t* vals_in = (int32_t*)(base + 2 * kb);
    void* tmp = base + 2 * kb + vb;
    k6_shuffle_keys_kernel<<<(nall + 255) / 256, 256, 0, st>>>(keys_in, vals_in, nall, seed, iteration, iter_dev);
    if (cub::DeviceRadixSort::SortPairs(tmp, cub_bytes, keys_in, keys_out, vals_in, perm, nall, 0, 64, st) != cudaSuccess)
        return JX_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}

static bool slice_ok(int nall, int split, int r_first, int r_count) {
    if (nall < 2 || (split != 0 && split != 1) || r_first < 0 || r_count < 0) return false;
    const int ns = (nall - split + 1) / 2;
    return r_first + r_count <= ns;
}

extern "C" int jx_stretch_propose(const double* coords, const int32_t* perm, int32_t nall, int32_t ndim,
                                  int32_t split, int32_t r_first, int32_t r_count, double a, uint64_t seed,
                                  uint64_t iteration, const uint64_t* iter_dev, double* prop, double* factor,
                                  int32_t device, void* stream) {
    if (!coords || !perm || !prop || !factor || ndim < 1 || !(a > 1.0)) return JX_ERR_INVALID;
    if (!slice_ok(nall, split, r_first, r_count)) return JX_ERR_INVALID;
    if (r_count == 0) return JX_OK;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_propose_kernel<<<(r_count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        coords, perm, nall, ndim, split, r_first, r_count, a, seed, iteration, iter_dev, prop, factor);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}

extern "C" int jx_stretch_accept(const double* coords, const double* lp, const int32_t* perm, int32_t nall,
                                 int32_t ndim, int32_t split, int32_t r_first, int32_t r_count, const double* prop,
                                 const double* lp_new, const double* factor, uint64_t seed, uint64_t iteration,
                                 const uint64_t* iter_dev, double* packed, int32_t device, void* stream) {
    if (!coords || !lp || !perm || !prop || !lp_new || !factor || !packed || ndim < 1) return JX_ERR_INVALID;
    if (!slice_ok(nall, split, r_first, r_count)) return JX_ERR_INVALID;
    if (r_count == 0) return JX_OK;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_accept_kernel<<<(r_count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        coords, lp, perm, ndim, split, r_first, r_count, prop, lp_new, factor, seed, iteration, iter_dev, packed);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}

extern "C" int jx_stretch_scatter(double* coords, double* lp, int32_t* naccept, const int32_t* perm, int32_t nall,
                                  int32_t ndim, int32_t split, const double* packed_all, int32_t ns, int32_t device,
                                  void* stream) {
    if (!coords || !lp || !perm || !packed_all || ndim < 1) return JX_ERR_INVALID;
    if (!slice_ok(nall, split, 0, ns)) return JX_ERR_INVALID;
    if (ns == 0) return JX_OK;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_scatter_kernel<<<(ns + 127) / 128, 128, 0, (cudaStream_t)stream>>>(coords, lp, naccept, perm, ndim, split,
                                                                            packed_all, ns);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}

extern "C" int jx_stretch_advance(uint64_t* iter_dev, uint64_t by, int32_t device, void* stream) {
    if (!iter_dev) return JX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(iter_dev, by);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}
