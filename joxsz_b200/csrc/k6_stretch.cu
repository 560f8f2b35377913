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

// ---- accept fused with the exchange: every rank writes its rows straight into the packed_all buffer of EVERY rank
// (peer memory over NVLink / NVSwitch), then raises one flag per peer; the scatter kernel of a rank waits for the flags
// of all ranks and reads its local copy.  No collective launch, no host involvement: the half-step's only
// communication is these stores.  Buffers are double-buffered by `split`: a rank can run at most one half-step ahead
// of a peer (its next scatter waits for that peer's accept), so the buffer of half-step k is never overwritten
// before every rank's scatter of k has read it.
constexpr int JX_MAX_PEERS = 16;
struct k6_peers {
    double* packed[JX_MAX_PEERS];     // [2][world * per0][ndim + 2] on every rank
    uint64_t* flags[JX_MAX_PEERS];    // [2][world] epochs on every rank
};

JX_D void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
JX_D uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

constexpr int K6P_THREADS = 128;

// The rows of a CTA (128 consecutive slice entries) are consecutive in every rank's buffer: they are staged in shared
// memory and then written to each peer as one contiguous block with coalesced 16-byte stores -- 8-byte stores scattered
// at the 120-byte row pitch cost one NVLink write per element and made this exchange slower than NCCL's (measured at 8
// ranks: 1.53 against 1.42 ms per iteration).
__global__ void __launch_bounds__(K6P_THREADS)
k6_accept_p2p_kernel(const double* __restrict__ coords, const double* __restrict__ lp,
                     const int32_t* __restrict__ perm, int ndim, int split, int r_first, int r_count,
                     const double* __restrict__ prop, const double* __restrict__ lp_new,
                     const double* __restrict__ factor, uint64_t seed, uint64_t iteration,
                     const uint64_t* __restrict__ iter_dev, k6_peers peers, int world, int rank, int per0,
                     unsigned int* __restrict__ done) {
    extern __shared__ __align__(16) double k6p_rows[];        // [K6P_THREADS][ndim + 2]
    const int i0 = blockIdx.x * K6P_THREADS, i = i0 + threadIdx.x;
    const int nrow = ndim + 2;
    iteration = effective_iteration(iteration, iter_dev);
    if (i < r_count) {
        const int k = perm[2 * (r_first + i) + split];
        philox4 r = philox4x32_10((uint32_t)k, (uint32_t)iteration, (uint32_t)(iteration >> 32),
                                  stream_word(PURPOSE_ACCEPT, split), (uint32_t)seed, (uint32_t)(seed >> 32));
        const double lnu = log(u01(r.v[0], r.v[1]));
        const double lnpdiff = factor[i] + lp_new[i] - lp[k];
        const bool acc = lnpdiff > lnu;
        const double* src = acc ? prop + (size_t)i * ndim : coords + (size_t)k * ndim;
        double* o = k6p_rows + (size_t)threadIdx.x * nrow;
        for (int d = 0; d < ndim; ++d) o[d] = src[d];
        o[ndim] = acc ? lp_new[i] : lp[k];
        o[ndim + 1] = acc ? 1.0 : 0.0;
    }
    __syncthreads();
    const int rows_here = min(K6P_THREADS, r_count - i0);                 // <= 0 for a CTA without rows
    if (rows_here > 0) {
        const size_t first = ((size_t)split * world * per0 + (size_t)rank * per0 + i0) * nrow;   // in doubles
        const int n = rows_here * nrow;
        for (int pp = 0; pp < world; ++pp) {
            const int p = (rank + 1 + pp) % world;                        // every rank starts at a different peer
            double* dst = peers.packed[p] + first;
            if (((first | (size_t)n) & 1) == 0) {                         // 16-byte aligned block: vector stores
                const double2* s2 = reinterpret_cast<const double2*>(k6p_rows);
                double2* d2 = reinterpret_cast<double2*>(dst);
                for (int j = threadIdx.x; j < n / 2; j += K6P_THREADS) d2[j] = s2[j];
            } else {
                for (int j = threadIdx.x; j < n; j += K6P_THREADS) dst[j] = k6p_rows[j];
            }
        }
    }
    __syncthreads();                           // the CTA's stores are ordered before thread 0's fence (cumulativity):
    if (threadIdx.x == 0) {                    // one system-scope fence per CTA instead of one per thread
        __threadfence_system();
        const unsigned int prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {           // last CTA of the grid: every row of this rank has been written
            *done = 0u;
            __threadfence_system();
            for (int p = 0; p < world; ++p)
                st_release_sys(peers.flags[p] + (size_t)split * world + rank, iteration + 1);
        }
    }
}

// rows of the half-step: r = 0..ns-1 in the order of the ranks' slices; rank g's slice starts at g * per in r and at
// g * per0 in the buffer (per <= per0: the second colour of an odd ensemble has one walker fewer)
__global__ void k6_scatter_p2p_kernel(double* __restrict__ coords, double* __restrict__ lp, int32_t* __restrict__ naccept,
                                      const int32_t* __restrict__ perm, int ndim, int split,
                                      const double* __restrict__ packed_local, const uint64_t* __restrict__ flags_local,
                                      int ns, int world, int per, int per0, uint64_t iteration,
                                      const uint64_t* __restrict__ iter_dev) {
    iteration = effective_iteration(iteration, iter_dev);
    if (threadIdx.x < world) {
        const uint64_t* f = flags_local + (size_t)split * world + threadIdx.x;
        while (ld_acquire_sys(f) < iteration + 1) __nanosleep(64);
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ns) return;
    const int g = r / per, off = r - g * per;
    const int k = perm[2 * r + split];
    const double* o = packed_local + ((size_t)split * world * per0 + (size_t)g * per0 + off) * (ndim + 2);
    for (int d = 0; d < ndim; ++d) coords[(size_t)k * ndim + d] = __ldcv(o + d);
    lp[k] = __ldcv(o + ndim);
    if (naccept && __ldcv(o + ndim + 1) != 0.0) naccept[k] += 1;
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

extern "C" int jx_stretch_accept_p2p(const double* coords, const double* lp, const int32_t* perm, int32_t nall,
                                     int32_t ndim, int32_t split, int32_t r_first, int32_t r_count, const double* prop,
                                     const double* lp_new, const double* factor, uint64_t seed, uint64_t iteration,
                                     const uint64_t* iter_dev, const uint64_t* peer_packed, const uint64_t* peer_flags,
                                     int32_t world, int32_t rank, int32_t per0, uint32_t* done, int32_t device,
                                     void* stream) {
    if (!coords || !lp || !perm || !peer_packed || !peer_flags || !done || ndim < 1) return JX_ERR_INVALID;
    if (world < 1 || world > JX_MAX_PEERS || rank < 0 || rank >= world || per0 < 1 || r_count > per0) return JX_ERR_INVALID;
    if (r_count > 0 && (!prop || !lp_new || !factor)) return JX_ERR_INVALID;
    if (!slice_ok(nall, split, r_first, r_count)) return JX_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_peers peers;
    for (int p = 0; p < JX_MAX_PEERS; ++p) {
        peers.packed[p] = p < world ? reinterpret_cast<double*>(peer_packed[p]) : nullptr;
        peers.flags[p] = p < world ? reinterpret_cast<uint64_t*>(peer_flags[p]) : nullptr;
        if (p < world && (!peers.packed[p] || !peers.flags[p])) return JX_ERR_INVALID;
    }
    const int grid = r_count > 0 ? (r_count + K6P_THREADS - 1) / K6P_THREADS : 1;   // a rank without rows still raises its flags
    const size_t smem = (size_t)K6P_THREADS * (ndim + 2) * sizeof(double);
    if (smem > 48 * 1024) return JX_ERR_INVALID;
    k6_accept_p2p_kernel<<<grid, K6P_THREADS, smem, (cudaStream_t)stream>>>(coords, lp, perm, ndim, split, r_first, r_count, prop,
                                                                 lp_new, factor, seed, iteration, iter_dev, peers, world,
                                                                 rank, per0, done);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}

extern "C" int jx_stretch_scatter_p2p(double* coords, double* lp, int32_t* naccept, const int32_t* perm, int32_t nall,
                                      int32_t ndim, int32_t split, const double* packed_local, const uint64_t* flags_local,
                                      int32_t ns, int32_t world, int32_t per, int32_t per0, uint64_t iteration,
                                      const uint64_t* iter_dev, int32_t device, void* stream) {
    if (!coords || !lp || !perm || !packed_local || !flags_local || ndim < 1 || split < 0 || split > 1) return JX_ERR_INVALID;
    if (world < 1 || world > JX_MAX_PEERS || per < 1 || per > per0) return JX_ERR_INVALID;
    if (ns != (nall - split + 1) / 2 || (size_t)per * world < (size_t)ns) return JX_ERR_INVALID;
    if (ns == 0) return JX_OK;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_scatter_p2p_kernel<<<(ns + 127) / 128, 128, 0, (cudaStream_t)stream>>>(coords, lp, naccept, perm, ndim, split,
                                                                               packed_local, flags_local, ns, world, per,
                                                                               per0, iteration, iter_dev);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}
