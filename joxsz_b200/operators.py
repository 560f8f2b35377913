"""Host-side construction of the constant operators the CUDA kernels consume.

Everything here runs ONCE, at set-up, in float64 numpy.  Nothing here evaluates a likelihood:
the outputs are constant tables (Abel projection matrix, spline-fit operators, beam spectrum,
cosine-transform matrices ...) that ``packer.py`` copies to the device through ``jx_create``.

Why fixed operators exist at all: every step of the reference's SZ chain between the pressure
profile and the filtered map row is *linear* with *fixed* geometry --

* ``abel.direct.direct_transform(pp, r=r_pp, 'forward')``  (reference ``joxsz_funcs.py:457``)
* ``interp1d(+-r_pp, (y, y), 'cubic')``                     (``joxsz_funcs.py:460``)
* ``fftconvolve(y_2d, beam_2d, 'same')``                    (``joxsz_funcs.py:464``)
* ``ifft2(fft2(conv) * filtering)``                         (``joxsz_funcs.py:466-467``)
* ``interp1d(radius[sep:], map_prof, 'cubic')``             (``joxsz_funcs.py:476``)

-- whereas the reference rebuilds the matrices behind each of them for every walker.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# Abel projection (replaces PyAbel's direct transform, Python backend)
# --------------------------------------------------------------------------------------


def is_uniform_sampling(r, atol=1e-13):
    """PyAbel's uniformity test: second differences of ``r`` all within ``atol`` of zero."""
    r = np.asarray(r, dtype=np.float64)
    return bool(np.allclose(np.diff(np.diff(r)), 0.0, atol=atol))


def abel_forward_matrix(r):
    """Matrix ``A`` with ``direct_transform(f, r=r, direction='forward') == A @ f``.

    Derivation from the algorithm of PyAbel's ``_pyabel_direct_integral``:
    row ``i`` integrates ``g_j = 2 r_j f_j / sqrt(r_j^2 - r_i^2)`` over ``j > i`` with the
    trapezoid rule on the full grid (``g = 0`` for ``j <= i``), removes half of the two
    trapezoids that touch the first sample ``j = i+1``, and adds the closed-form integral of the
    singular cell ``[r_i, r_{i+1}]`` for a piecewise-linear integrand.  The factor ``2 r_j`` of the
    forward transform is folded in.  The last row is zero.
    """
    r = np.asarray(r, dtype=np.float64)
    n = r.size
    if n < 3:
        raise ValueError("need at least 3 radial points")
    dr = np.diff(r)
    uniform = is_uniform_sampling(r)
    # trapezoid node weights for a function sampled on the whole grid
    if uniform:
        dx = abs(r[1] - r[0])
        wl = np.full(n - 1, 0.5 * dx)
    else:
        wl = 0.5 * dr
    node_w = np.zeros(n)
    node_w[:-1] += wl
    node_w[1:] += wl

    A = np.zeros((n, n))
    ri = r[:, None]
    rj = r[None, :]
    upper = np.triu(np.ones((n, n), dtype=bool), k=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        root = np.sqrt(np.where(upper, rj * rj - ri * ri, 1.0))
    inv_root = np.where(upper, 1.0 / root, 0.0)
    A += inv_root * node_w[None, :]
    # "extra triangle": half of the trapezoid integral of the single spike at j = i+1
    idx = np.arange(n - 1)
    A[idx, idx + 1] -= 0.5 * node_w[idx + 1] * inv_root[idx, idx + 1]
    # analytic first cell, f piecewise linear: s*f' + acosh(r1/r0)*(f0 - f' r0)
    s = root[idx, idx + 1]
    if r[0] < r[1] * 1e-8:
        ratio = np.append(np.cosh(1.0), r[2:] / r[1:-1])
    else:
        ratio = r[1:] / r[:-1]
    acr = np.arccosh(ratio)
    A[idx, idx] += acr * (1.0 + r[:-1] / dr) - s / dr
    A[idx, idx + 1] += s / dr - acr * r[:-1] / dr
    return A * (2.0 * r)[None, :]


# --------------------------------------------------------------------------------------
# Not-a-knot cubic splines as linear operators
# --------------------------------------------------------------------------------------


def notaknot_second_derivative_operator(x):
    """``Minv`` [n, n] with ``M = Minv @ y``: knot second derivatives of the not-a-knot cubic
    interpolating spline through ``(x, y)`` -- what ``scipy.interpolate.interp1d(kind='cubic')``
    (= ``make_interp_spline(k=3)``) constructs."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 4:
        raise ValueError("not-a-knot cubic needs >= 4 knots")
    if not np.all(np.diff(x) > 0):
        raise ValueError("knots must be strictly increasing")
    h = np.diff(x)
    T = np.zeros((n, n))
    R = np.zeros((n, n))
    for i in range(1, n - 1):
        T[i, i - 1] = h[i - 1]
        T[i, i] = 2.0 * (h[i - 1] + h[i])
        T[i, i + 1] = h[i]
        R[i, i - 1] = 6.0 / h[i - 1]
        R[i, i] = -6.0 / h[i - 1] - 6.0 / h[i]
        R[i, i + 1] = 6.0 / h[i]
    # third derivative continuous across x[1] and x[n-2]
    T[0, 0], T[0, 1], T[0, 2] = h[1], -(h[0] + h[1]), h[0]
    T[n - 1, n - 3], T[n - 1, n - 2], T[n - 1, n - 1] = h[n - 2], -(h[n - 3] + h[n - 2]), h[n - 3]
    return np.linalg.solve(T, R)


def spline_piece_operators(x):
    """Per-interval polynomial coefficients as linear maps of the data.

    Returns ``C`` [4, n-1, n]; on ``[x_k, x_{k+1}]`` the spline is
    ``sum_p (C[p, k] @ y) * (t - x_k)**p``.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    h = np.diff(x)
    Minv = notaknot_second_derivative_operator(x)
    eye = np.eye(n)
    C = np.empty((4, n - 1, n))
    C[0] = eye[:-1]
    C[1] = (eye[1:] - eye[:-1]) / h[:, None] - h[:, None] * (2.0 * Minv[:-1] + Minv[1:]) / 6.0
    C[2] = 0.5 * Minv[:-1]
    C[3] = (Minv[1:] - Minv[:-1]) / (6.0 * h[:, None])
    return C


def spline_eval_operator(x, xq, extrapolate=True):
    """``E`` [len(xq), n] with ``E @ y`` = not-a-knot cubic spline through ``(x, y)`` at ``xq``.

    Outside ``[x0, xn]`` the end polynomials are continued (``fill_value='extrapolate'``).
    """
    x = np.asarray(x, dtype=np.float64)
    xq = np.atleast_1d(np.asarray(xq, dtype=np.float64))
    C = spline_piece_operators(x)
    k = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, x.size - 2)
    if not extrapolate and (np.any(xq < x[0]) or np.any(xq > x[-1])):
        raise ValueError("query outside the knot range")
    t = xq - x[k]
    return C[0, k] + t[:, None] * (C[1, k] + t[:, None] * (C[2, k] + t[:, None] * C[3, k]))


def symmetric_knots(r):
    """Sorted knot vector of ``interp1d(np.append(-r, r), ...)`` (interp1d sorts its abscissae)."""
    r = np.asarray(r, dtype=np.float64)
    return np.concatenate((-r[::-1], r))


def symmetric_fold(n):
    """``F`` [2n, n]: values ``(y, y)`` on the sorted knots ``(-r[::-1], r)`` from ``y`` on ``r``."""
    eye = np.eye(n)
    return np.concatenate((eye[::-1], eye), axis=0)


# --------------------------------------------------------------------------------------
# Geometry checks for the quarter-plane (D4-symmetric) map pipeline
# --------------------------------------------------------------------------------------


class GeometryError(ValueError):
    """The SZ_data geometry is not one the CUDA map kernel supports; raised at pack time."""


def _check_d4(name, a, rtol=1e-12):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] != a.shape[1] or a.shape[0] % 2 != 1:
        raise GeometryError(f"{name}: expected an odd-sized square image, got {a.shape} "
                            "(the reference builds maps of side 2m+1, joxsz_main.py:101-103)")
    scale = np.max(np.abs(a)) or 1.0
    for what, b in (("left-right", a[:, ::-1]), ("up-down", a[::-1, :]), ("transpose", a.T)):
        if np.max(np.abs(a - b)) > rtol * scale:
            raise GeometryError(f"{name}: not {what} symmetric about its centre pixel")


def quarter(a):
    """Lower-right quadrant including the centre row/column of an odd-sized image."""
    c = a.shape[0] // 2
    return np.ascontiguousarray(a[c:, c:])


SUPPORTED_FFT_SIZES = (256, 512, 1024)     # 256: shared-memory map kernel; 512 / 1024: L2-staged map kernel


def choose_padded_size(n_map, n_beam, supported=SUPPORTED_FFT_SIZES):
    """Smallest supported cyclic length that reproduces ``fftconvolve(..., 'same')`` exactly.

    With beam half-width ``b`` the cropped output uses full-convolution indices ``b .. b+N-1``;
    wrap-around of a length-``P`` cyclic convolution pollutes indices ``0 .. N+B-2-P``, so any
    ``P >= N + b`` is exact (scipy itself picks ``next_fast_len(N+B-1)``).
    """
    need = n_map + n_beam // 2
    for p in supported:
        if p >= need:
            return p
    raise GeometryError(f"map side {n_map} with beam side {n_beam} needs a cyclic length >= {need}; "
                        f"supported by the CUDA FFT: {supported}")


def _cos_table(nrow, ncol, period):
    """cos(2 pi i j / period) with the integer product reduced mod ``period`` first."""
    i = np.arange(nrow, dtype=np.int64)[:, None]
    j = np.arange(ncol, dtype=np.int64)[None, :]
    return np.cos(2.0 * np.pi * ((i * j) % period) / period)


def _fold_weights(n):
    w = np.full(n, 2.0)
    w[0] = 1.0
    return w


class SZMapOperators:
    """Constant tables of the map stage (kernel K3) for one SZ_data geometry.

    Notation: ``N`` map side (odd), ``c = N//2``, ``H = c+1`` quarter-plane side, ``B`` beam
    side (odd), ``P`` cyclic FFT length, ``Q = P/2+1``.  All 2-D arrays of the stage (Compton-y
    map, beam, convolved map, filter) are symmetric under x->-x, y->-y about the centre pixel, so
    re-centred on the origin their DFTs are real cosine transforms of the quarter plane.

    Tables (all float64, C order):

    ``seg`` [H,H] int32, ``dx`` [H,H]  spline piece index (0 = the central piece ``[-r_1, r_1]``,
                                       k = ``[r_k, r_{k+1}]``) and offset from the piece's left knot
                                       for the map pixel at centred offset (u, v)
    ``nseg``                           number of pieces referenced (pieces 0..nseg-1)
    ``bhat`` [Q,Q]                     ``step^2 / P^2 * sum_{u,v} beam_c[u,v] cos(2pi ky u/P) cos(2pi kx v/P)``
    ``bmix`` [B//2+1,Q]                ``step^2 / P * sum_v beam_c[j,v] cos(2pi kx v/P)``: beam in (y offset, kx)
    ``cmat`` [H,H]                     ``w_v cos(2 pi kx v / N)`` indexed [v, kx]: row DCT at length N
    ``hf``   [H,H]                     ``w_u sum_ky filt[ky,kx] cos(2 pi ky u / N)`` indexed [u, kx]
    ``dinv`` [H,H]                     ``w_kx cos(2 pi kx v / N) / N^2`` indexed [kx, v]
    """

    def __init__(self, r_pp, d_mat, beam_2d, filtering, step):
        r_pp = np.asarray(r_pp, dtype=np.float64)
        d_mat = np.asarray(d_mat, dtype=np.float64)
        beam_2d = np.asarray(beam_2d, dtype=np.float64)
        filtering = np.asarray(filtering, dtype=np.float64)
        _check_d4("d_mat", d_mat)
        _check_d4("beam_2d", beam_2d)
        if filtering.shape != d_mat.shape:
            raise GeometryError("filtering must have the shape of d_mat (joxsz_main.py:107)")
        N = d_mat.shape[0]
        B = beam_2d.shape[0]
        c = N // 2
        H = c + 1
        # filtering lives in FFT order: index k and N-k are the same |k|
        f_ref = filtering[(-np.arange(N)) % N][:, (-np.arange(N)) % N]
        scale = np.max(np.abs(filtering)) or 1.0
        if (np.max(np.abs(filtering - f_ref)) > 1e-12 * scale
                or np.max(np.abs(filtering - filtering[(-np.arange(N)) % N])) > 1e-12 * scale
                or np.max(np.abs(filtering - filtering.T)) > 1e-12 * scale):
            raise GeometryError("filtering is not an even function of (kx, ky)")
        if d_mat[c, c] != 0.0:
            raise GeometryError("d_mat centre pixel must be at distance 0")
        self.N, self.B, self.H, self.c = N, B, H, c
        self.P = choose_padded_size(N, B)
        self.Q = self.P // 2 + 1
        P, Q = self.P, self.Q

        # --- map synthesis: which spline piece each quarter-plane pixel falls in
        dq = quarter(d_mat)
        if dq.max() > r_pp[-1]:
            raise GeometryError("map extends beyond r_pp[-1]: the reference's fill_value=(0,0) branch "
                                "(joxsz_funcs.py:460) is not implemented")
        knots = symmetric_knots(r_pp)
        k_abs = np.clip(np.searchsorted(knots, dq, side="right") - 1, 0, knots.size - 2)
        k_rel = k_abs - (r_pp.size - 1)          # 0 = central piece [-r_1, +r_1]
        if k_rel.min() < 0:
            raise GeometryError("negative distance in d_mat")
        self.seg = k_rel.astype(np.int32)
        self.dx = dq - knots[k_abs]
        self.nseg = int(k_rel.max()) + 1

        # --- beam spectrum on the quarter plane of the length-P cyclic grid
        b = B // 2
        bq = quarter(beam_2d)                     # [b+1, b+1], centred offsets 0..b
        wb = _fold_weights(b + 1)
        cb = _cos_table(Q, b + 1, P) * wb[None, :]   # [k, u]
        self.bhat = (cb @ bq @ cb.T) * (float(step) ** 2 / float(P) ** 2)
        # mixed domain (pixel offset j along y, frequency kx along x): the beam convolution along y done directly,
        # conv[u, kx] = sum_j bmix[|j|, kx] ext(X1)[u - j, kx]; carries step^2 and the 1/P of the x transforms
        self.bmix = np.ascontiguousarray((bq @ cb.T) * (float(step) ** 2 / float(P)))   # [b+1, Q]

        # --- exact length-N circular filter, reduced to the one row that is consumed
        wN = _fold_weights(H)
        cosN = _cos_table(H, H, N)                # [i, j] = cos(2 pi i j / N)
        self.cmat = np.ascontiguousarray(cosN * wN[:, None])          # [v, kx]
        fq = filtering[:H, :H]                    # |ky|, |kx| = 0..c
        # sum over all ky in 0..N-1 of filt[ky,kx] cos(2 pi ky u/N) = sum_{ky<=c} w_ky filt cos
        self.hf = np.ascontiguousarray(((cosN * wN[None, :]) @ fq) * wN[:, None])  # [u, kx]
        self.dinv = np.ascontiguousarray(cosN * wN[:, None] / float(N) ** 2)       # [kx, v]
        self.filt_q = np.ascontiguousarray(fq)
        self.cosN = cosN
        self.wN = wN


def sz_spline_coeff_operator(r_pp, nseg):
    """``G`` [4*nseg, Nr]: polynomial coefficients of pieces 0..nseg-1 of the cubic spline through
    ``(+-r_pp, (f, f))`` as a linear map of ``f`` on ``r_pp`` (row ``p*nseg + k`` = coefficient
    ``p`` of piece ``k``, local variable ``t - left_knot``)."""
    r_pp = np.asarray(r_pp, dtype=np.float64)
    nr = r_pp.size
    C = spline_piece_operators(symmetric_knots(r_pp))      # [4, 2nr-1, 2nr]
    F = symmetric_fold(nr)                                  # [2nr, nr]
    pieces = C[:, nr - 1:nr - 1 + nseg, :] @ F              # [4, nseg, nr]
    return pieces.reshape(4 * nseg, nr)


def central_value_operator(r):
    """Weights ``w`` with ``w @ t`` = value at 0 of the cubic spline through ``(+-r, (t, t))``
    (``h(0.)``, reference ``joxsz_funcs.py:470-473``)."""
    r = np.asarray(r, dtype=np.float64)
    E = spline_eval_operator(symmetric_knots(r), np.array([0.0]))
    return (E @ symmetric_fold(r.size))[0]
