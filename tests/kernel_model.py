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


def model_stages(pk, pp):
    """pk: PackedSetup, pp [W, nr] -> dict of stage outputs as the kernels define them."""
    mo = pk.map_ops
    W = pp.shape[0]
    H, P, Q, N = mo.H, mo.P, mo.Q, mo.N
    coef = (pp @ pk.proj_op.T).reshape(W, 4, mo.nseg)
    seg, dx = pk.seg, pk.dx
    Z = coef[:, 0][:, seg] + dx * (coef[:, 1][:, seg] + dx * (coef[:, 2][:, seg] + dx * coef[:, 3][:, seg]))
    X1 = dct_even(Z, P, Q)                                   # [W, H(u), Q(kx)]   phase A
    S = dct_even(np.swapaxes(X1, 1, 2), P, Q)                # [W, Q(kx), Q(ky)]  phase B forward
    S = S * pk.bhat.T[None]                                  # bhat[ky, kx]
    X2 = np.swapaxes(dct_even(S, P, H), 1, 2)                # [W, H(u), Q(kx)]   phase B inverse
    conv = dct_even(X2, P, H)                                # [W, H(u), H(v)]    phase C
    C1 = conv @ pk.cmat                                      # [W, u, kx]         phase D
    G = np.einsum("wuk,uk->wk", C1, pk.hf)
    row = G @ pk.dinv                                        # [W, v]             phase E
    return dict(coef=coef, Z=Z, conv=conv, row=row)


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
