"""numpy model of what the CUDA kernels K2/K3/K5 compute from the packed constant tables.

Test infrastructure: it re-derives, with numpy FFTs and the tables of joxsz_b200.packer.PackedSetup,
exactly the sequence of operations the map kernel performs (even extensions, cosine transforms of the
quarter plane, packed beam spectrum, dense length-N cosine transform reduced over ky).  It validates
the *tables and the algorithm* on the CPU, so that what remains to check on the GPU is the CUDA code.
"""
import numpy as np


def even_ext(a, P):
    """Length-P even extension along the last axis of a half-array a[..., 0:m] (zero beyond m-1)."""
    m = a.shape[-1]
    out = np.zeros(a.shape[:-1] + (P,))
    out[..., :m] = a
    out[..., P - (m - 1):] = a[..., :0:-1] if m > 1 else 0
    return out


def dct_even(a, P, nout):
    """sum_n ext(a)[n] cos(2 pi n k / P), k < nout, along the last axis."""
    return np.real(np.fft.fft(even_ext(a, P), axis=-1))[..., :nout]


def model_stages(pk, pp, direct_b=False):
    """pk: PackedSetup, pp [W, nr] -> dict of stage outputs as the kernels define them."""
    mo = pk.map_ops
    W = pp.shape[0]
    H, P, Q, N = mo.H, mo.P, mo.Q, mo.N
    coef = (pp @ pk.proj_op.T).reshape(W, 4, mo.nseg)
    seg, dx = pk.seg, pk.dx
    Z = coef[:, 0][:, seg] + dx * (coef[:, 1][:, seg] + dx * (coef[:, 2][:, seg] + dx * coef[:, 3][:, seg]))
    X1 = dct_even(Z, P, Q)                                   # [W, H(u), Q(kx)]   phase A
    if direct_b:
        # phase B as the kernel does it when the beam is small: direct convolution along y in the mixed domain
        nb = pk.bmix.shape[0]
        ext = np.zeros((W, H + 2 * (nb - 1), Q))             # rows -(nb-1) .. H+nb-2 of the even extension
        ext[:, nb - 1:nb - 1 + H] = X1
        ext[:, :nb - 1] = X1[:, nb - 1:0:-1]
        X2 = np.zeros_like(X1)
        for j in range(-(nb - 1), nb):
            X2 += pk.bmix[abs(j)][None, None, :] * ext[:, nb - 1 - j:nb - 1 - j + H]
    else:
        S = dct_even(np.swapaxes(X1, 1, 2), P, Q)            # [W, Q(kx), Q(ky)]  phase B forward
        S = S * pk.bhat.T[None]                              # bhat[ky, kx]
        X2 = np.swapaxes(dct_even(S, P, H), 1, 2)            # [W, H(u), Q(kx)]   phase B inverse
    conv = dct_even(X2, P, H)                                # [W, H(u), H(v)]    phase C
    # K7: packed triangle u <= v of the convolved map times the filter-row operator
    iu, iv = np.triu_indices(H)
    tri = conv[:, iu, iv]                                    # [W, H (H + 1) / 2], row-major packed
    if H <= 136:
        row = tri @ filter_row_operator(pk)                  # [W, v]
    else:
        # wide quarter planes: the same contraction without materialising the 32 896 x 256 operator in long double
        C1 = conv @ pk.cmat                                  # [W, u, kx]
        row = np.einsum("wuk,uk->wk", C1, pk.hf) @ pk.dinv   # [W, v]
    return dict(coef=coef, Z=Z, conv=conv, row=row)


def filter_row_operator(pk):
    """R [H (H + 1) / 2, H] as jx_create builds it for K7 (csrc/jx_api.cu): response of map_out[N//2, N//2 + x] to the
    convolved-map pixel pair conv_c[u, v] = conv_c[v, u], u <= v row-major."""
    H = pk.map_ops.H
    iu, iv = np.triu_indices(H)
    hf, cmat, dinv = (np.asarray(a, dtype=np.longdouble) for a in (pk.hf, pk.cmat, pk.dinv))
    F = hf[iu] * cmat[iv] + np.where((iu != iv)[:, None], hf[iv] * cmat[iu], 0.0)      # [(u,v), kx]
    return np.asarray(F @ dinv, dtype=np.float64)


def model_tail(pk, row, tsz, calib):
    """phase F: brightness, model at the data radii, chi^2."""
    T0 = tsz @ pk.w_t0
    T = np.concatenate([T0[:, None], tsz], axis=1)
    xk, yk = pk.conv_T, pk.conv_I
    idx = np.clip(np.searchsorted(xk, T), 1, xk.size - 1)
    slope = (yk[idx] - yk[idx - 1]) / (xk[idx] - xk[idx - 1])
    conv = slope * (T - xk[idx - 1]) + yk[idx - 1]
    bright = row * conv * calib[:, None]
    model = bright @ pk.g_op.T
    z = ((pk.flux - model) / pk.flux_err) ** 2
    return dict(bright=bright, model=model, chisq=np.nansum(z, axis=1))


# ---------------------------------------------------------------------------------------------
# K6: numpy model of the stretch-move kernels (Philox4x32-10 counter RNG) -- lets the sampler's
# sharding / index logic run on CPU (gloo, world_size 2) and pins the CUDA kernels' random streams.
# ---------------------------------------------------------------------------------------------

def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments uint32 arrays (broadcastable). Returns 4 uint32 arrays."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c0, c1, c2, c3, k0, k1 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3, k0, k1))
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    c0, c1, c2, c3, k0, k1 = (v.copy() for v in (c0, c1, c2, c3, k0, k1))
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = p0.astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = k0 + W0
            k1 = k1 + W1
    return c0, c1, c2, c3


def _u01(hi, lo):
    x = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _draws(k, seed, iteration, split, purpose):
    seed, iteration = int(seed), int(iteration)
    return philox4x32_10(k.astype(np.uint32), np.uint32(iteration & 0xFFFFFFFF), np.uint32(iteration >> 32),
                         np.uint32(purpose | (split << 2)), np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32))


def split_permutation(nall, seed, iteration):
    """Model of jx_stretch_permutation: stable argsort of 64 Philox bits per walker.  The colour of the walker
    at position p is p & 1, i.e. emcee's ``inds = arange(n) % 2; random.shuffle(inds)``."""
    r0, r1, _, _ = _draws(np.arange(nall), seed, iteration, 0, 2)      # purpose 2 = colouring keys (own Philox block)
    keys = (r0.astype(np.uint64) << np.uint64(32)) | r1.astype(np.uint64)
    return np.argsort(keys, kind="stable").astype(np.int32)


def stretch_propose(coords, perm, split, r_first, r_count, a, seed, iteration):
    nall, ndim = coords.shape
    i = np.arange(r_count)
    k = perm[2 * (r_first + i) + split]
    r0, r1, r2, _ = _draws(k, seed, iteration, split, 0)
    u = _u01(r0, r1)
    root = (a - 1.0) * u + 1.0
    z = root * root / a
    other = 1 - split
    nc = (nall - other + 1) // 2
    rint = ((r2.astype(np.uint64) * np.uint64(nc)) >> np.uint64(32)).astype(np.int64)
    s = coords[k]
    c = coords[perm[2 * rint + other]]
    prop = c - (c - s) * z[:, None]
    factor = (ndim - 1.0) * np.log(z)
    return prop, factor


def stretch_accept(coords, lp, perm, split, r_first, r_count, prop, lp_new, factor, seed, iteration):
    ndim = coords.shape[1]
    i = np.arange(r_count)
    k = perm[2 * (r_first + i) + split]
    r0, r1, _, _ = _draws(k, seed, iteration, split, 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        lnu = np.log(_u01(r0, r1))
        acc = (factor + lp_new - lp[k]) > lnu
    packed = np.empty((r_count, ndim + 2))
    packed[:, :ndim] = np.where(acc[:, None], prop, coords[k])
    packed[:, ndim] = np.where(acc, lp_new, lp[k])
    packed[:, ndim + 1] = acc.astype(np.float64)
    return packed


def stretch_scatter(coords, lp, naccept, perm, split, packed_all, ns):
    ndim = coords.shape[1]
    k = perm[2 * np.arange(ns) + split]
    coords[k] = packed_all[:ns, :ndim]
    lp[k] = packed_all[:ns, ndim]
    naccept[k] += (packed_all[:ns, ndim + 1] != 0).astype(naccept.dtype)


class NumpyStretchOps:
    """Same interface as joxsz_b200.sampler.CudaStretchOps, on CPU torch tensors (tests only)."""
    launches_per_half_step = 3

    def permutation(self, perm, seed, iteration):
        perm.numpy()[:] = split_permutation(perm.shape[0], seed, iteration)

    def propose(self, coords, perm, split, r_first, r_count, a, seed, iteration, prop, factor):
        p, f = stretch_propose(coords.numpy(), perm.numpy(), split, r_first, r_count, a, seed, iteration)
        prop.numpy()[:r_count] = p
        factor.numpy()[:r_count] = f

    def accept(self, coords, lp, perm, split, r_first, r_count, prop, lp_new, factor, seed, iteration, packed):
        packed.numpy()[:r_count] = stretch_accept(coords.numpy(), lp.numpy(), perm.numpy(), split, r_first, r_count,
                                                  prop.numpy()[:r_count], lp_new.numpy()[:r_count],
                                                  factor.numpy()[:r_count], seed, iteration)

    def scatter(self, coords, lp, naccept, perm, split, packed_all, ns):
        stretch_scatter(coords.numpy(), lp.numpy(), naccept.numpy(), perm.numpy(), split, packed_all.numpy(), ns)


class GaussianToyLikelihood:
    """log N(0, diag(sig^2)) on CPU torch tensors with the engine interface the sampler needs."""

    def __init__(self, ndim, max_walkers=1 << 20):
        import torch
        self.ndim, self.max_walkers = ndim, max_walkers
        self.device = torch.device("cpu")
        self.sig = torch.linspace(0.5, 2.0, ndim, dtype=torch.float64)
        self.calls = []

    def loglike_device(self, theta, out=None):
        ll = -0.5 * ((theta / self.sig) ** 2).sum(dim=1)
        self.calls.append(theta.shape[0])
        if out is not None:
            out.copy_(ll)
            return out
        return ll

    def __call__(self, theta):
        import torch
        return self.loglike_device(torch.from_numpy(np.atleast_2d(np.asarray(theta, float)))).numpy()
