// K6 -- affine-invariant stretch move (Goodman & Weare 2010) with emcee's red/blue split.
//
// Replaces, for the walkers owned by this rank, emcee's StretchMove.get_proposal and the accept loop
// of RedBlueMove.propose (emcee 3.x; driven by the reference at joxsz_funcs.py:593-622 through
// EnsembleSampler, joxsz_main.py:206-210).  The likelihood of the proposals is evaluated in between
// by jx_loglike.  Random numbers come from Philox4x32-10 keyed by the run seed with the counter
// (global walker index, iteration, split | purpose), so a chain is bit-identical however the
// ensemble is sharded over GPUs.
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

constexpr uint32_t PURPOSE_PROPOSE = 0u, PURPOSE_ACCEPT = 1u;

__global__ void k6_propose_kernel(const double* __restrict__ coords, const int32_t* __restrict__ perm,
                                  const int32_t* __restrict__ pos, int nall, int ndim, int first, int count,
                                  int split, double a, uint64_t seed, uint64_t iteration, double* __restrict__ prop,
                                  double* __restrict__ factor, int32_t* __restrict__ active) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int gi = first + i;
    const double* s = coords + (size_t)gi * ndim;
    double* q = prop + (size_t)i * ndim;
    if ((pos[gi] & 1) != split) {
        for (int k = 0; k < ndim; ++k) q[k] = s[k];
        factor[i] = 0.0;
        active[i] = 0;
        return;
    }
    philox4 r = philox4x32_10((uint32_t)gi, (uint32_t)iteration, (uint32_t)(iteration >> 32),
                              ((uint32_t)split << 1) | PURPOSE_PROPOSE, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = u01(r.v[0], r.v[1]);
    const double root = (a - 1.0) * u + 1.0;
    const double z = root * root / a;
    const int other = 1 - split;
    const int nc = (nall - other + 1) / 2;                // positions of parity `other` in 0..nall-1
    int rint = (int)(((uint64_t)r.v[2] * (uint64_t)nc) >> 32);
    const double* c = coords + (size_t)perm[2 * rint + other] * ndim;
    for (int k = 0; k < ndim; ++k) q[k] = c[k] - (c[k] - s[k]) * z;
    factor[i] = ((double)ndim - 1.0) * log(z);
    active[i] = 1;
}

__global__ void k6_accept_kernel(double* __restrict__ coords_local, double* __restrict__ lp_local,
                                 const double* __restrict__ prop, const double* __restrict__ lp_new,
                                 const double* __restrict__ factor, const int32_t* __restrict__ active, int ndim,
                                 int first, int count, int split, uint64_t seed, uint64_t iteration,
                                 int32_t* __restrict__ naccept) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count || !active[i]) return;
    const int gi = first + i;
    philox4 r = philox4x32_10((uint32_t)gi, (uint32_t)iteration, (uint32_t)(iteration >> 32),
                              ((uint32_t)split << 1) | PURPOSE_ACCEPT, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double lnu = log(u01(r.v[0], r.v[1]));
    const double lnpdiff = factor[i] + lp_new[i] - lp_local[i];
    if (lnpdiff > lnu) {
        for (int k = 0; k < ndim; ++k) coords_local[(size_t)i * ndim + k] = prop[(size_t)i * ndim + k];
        lp_local[i] = lp_new[i];
        if (naccept) naccept[i] += 1;
    }
}

}  // namespace

extern "C" int jx_stretch_propose(const double* coords, const int32_t* perm, const int32_t* pos, int32_t nall,
                                  int32_t ndim, int32_t first, int32_t count, int32_t split, double a,
                                  uint64_t seed, uint64_t iteration, double* prop, double* factor,
                                  int32_t* active, int32_t device, void* stream) {
    if (!coords || !perm || !pos || !prop || !factor || !active) return JX_ERR_INVALID;
    if (nall < 2 || ndim < 1 || first < 0 || count < 0 || first + count > nall || (split != 0 && split != 1) ||
        !(a > 1.0))
        return JX_ERR_INVALID;
    if (count == 0) return JX_OK;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_propose_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        coords, perm, pos, nall, ndim, first, count, split, a, seed, iteration, prop, factor, active);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}

extern "C" int jx_stretch_accept(double* coords_local, double* lp_local, const double* prop, const double* lp_new,
                                 const double* factor, const int32_t* active, int32_t ndim, int32_t first,
                                 int32_t count, int32_t split, uint64_t seed, uint64_t iteration, int32_t* naccept,
                                 int32_t device, void* stream) {
    if (!coords_local || !lp_local || !prop || !lp_new || !factor || !active) return JX_ERR_INVALID;
    if (ndim < 1 || first < 0 || count < 0 || (split != 0 && split != 1)) return JX_ERR_INVALID;
    if (count == 0) return JX_OK;
    if (cudaSetDevice(device) != cudaSuccess) return JX_ERR_CUDA;
    k6_accept_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        coords_local, lp_local, prop, lp_new, factor, active, ndim, first, count, split, seed, iteration, naccept);
    return cudaGetLastError() == cudaSuccess ? JX_OK : JX_ERR_CUDA;
}
