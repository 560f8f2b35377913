"""Host-side readers and SZ set-up helpers with the reference's names and argument meaning.

These mirror ``joxsz_funcs.py:16-211`` (NIKA beam / transfer function / flux profile readers,
beam image, distance matrix, transfer-function image, ``SZ_data`` bag, Chandra band loaders).
They run once, before sampling; the likelihood kernels only ever see the arrays they return.
FITS tables are decoded by :mod:`joxsz_b200.fitsio` because astropy is not a dependency here.
"""
from __future__ import annotations

import os

import numpy as np
from scipy import optimize
from scipy.interpolate import interp1d
from scipy.stats import norm

from . import fitsio
from .mb import mb


def read_xy_err(filename, ncol):
    """First ``ncol`` columns of a FITS bintable (row 0 of the first extension) or of an ASCII
    table (``.txt``/``.dat``).  Same contract as reference ``joxsz_funcs.py:16-28``."""
    ext = os.path.splitext(filename)[1].lower().lstrip(".")
    if ext == "fits":
        cols = fitsio.read_bintable(filename, ext=1)[0]
    elif ext in ("txt", "dat"):
        cols = np.loadtxt(filename, unpack=True)
    else:
        raise RuntimeError("Unrecognised file extension (not in fits, dat, txt)")
    return cols[:ncol]


def read_beam(filename):
    """Beam radius/profile, truncated before the first NaN and then before the first negative
    sample (reference ``joxsz_funcs.py:30-44``)."""
    radius, prof = (np.asarray(a, dtype=np.float64) for a in read_xy_err(filename, ncol=2))
    for bad in (np.isnan(prof), prof < 0.0):
        hit = np.flatnonzero(bad)
        if hit.size:
            radius, prof = radius[:hit[0]], prof[:hit[0]]
    return radius, prof


def centdistmat(r, offset=0.0):
    """Matrix of distances from the centre for a symmetric axis vector ``r``
    (reference ``joxsz_funcs.py:78-88``)."""
    r = np.asarray(r, dtype=np.float64)
    return np.sqrt(r[None, :] ** 2 + r[:, None] ** 2) + offset


def beam_image(step, maxr_data, r_tab=None, b_tab=None, approx=False, normalize=True, fwhm_beam=None):
    """Beam image from an already-loaded radial table (or a Gaussian when ``approx``); the
    file-free core of :func:`mybeam`."""
    profile = None
    if not approx:
        profile = interp1d(np.concatenate((-r_tab, r_tab)), np.concatenate((b_tab, b_tab)), "cubic",
                           bounds_error=False, fill_value=(0.0, 0.0))
        half = profile(0.0) / 2
        fwhm_beam = 2 * optimize.newton(lambda x: profile(x) - half, x0=5.0)
    maxr = (maxr_data + 3 * fwhm_beam) // step * step
    pos = np.arange(0.0, maxr + step, step)
    axis = np.concatenate((-pos[:0:-1], pos))
    axis = axis[np.abs(axis) <= 3 * fwhm_beam]
    dist_img = centdistmat(axis)
    if approx:
        sigma = fwhm_beam / (2 * np.sqrt(2 * np.log(2)))
        beam_2d = norm.pdf(dist_img, loc=0.0, scale=sigma)
    else:
        beam_2d = profile(dist_img)
    if normalize:
        beam_2d = beam_2d / (beam_2d.sum() * step ** 2)
    return beam_2d, fwhm_beam


def mybeam(step, maxr_data, approx=False, filename=None, normalize=True, fwhm_beam=None):
    """2-D beam image and its FWHM (reference ``joxsz_funcs.py:46-76``).

    From file: the tabulated profile is mirrored, cubic-interpolated (zero outside the table),
    its FWHM found by Newton iteration on ``f(x) - f(0)/2`` from x0 = 5, and the image sampled on
    a grid of pitch ``step`` cut at 3 FWHM.  ``approx=True`` uses a Gaussian of the given FWHM.
    """
    r_tab = b_tab = None
    if not approx:
        r_tab, b_tab = read_beam(filename)
    return beam_image(step, maxr_data, r_tab, b_tab, approx=approx, normalize=normalize, fwhm_beam=fwhm_beam)


def read_tf(filename, approx=False, loc=0.0, scale=0.02, c=0.95):
    """Wave numbers (1/arcsec) and transmission; optionally a normal-cdf approximation
    (reference ``joxsz_funcs.py:90-102``)."""
    wn, tf = (np.asarray(a, dtype=np.float64) for a in read_xy_err(filename, ncol=2))
    if approx:
        tf = c * norm.cdf(wn, loc, scale)
    return wn, tf


def dist(naxis):
    """IDL ``DIST``-like matrix: each element proportional to its FFT frequency
    (reference ``joxsz_funcs.py:104-116``)."""
    axis = np.linspace(-naxis // 2 + 1, naxis // 2, naxis)
    grid = np.sqrt(axis[None, :] ** 2 + axis[:, None] ** 2)
    return np.roll(grid, naxis // 2 + 1, axis=(0, 1))


def filt_image(wn_as, tf, side, step):
    """``side x side`` transfer-function image in FFT order (reference ``joxsz_funcs.py:118-134``)."""
    tf_of_k = interp1d(wn_as, tf, "cubic", bounds_error=False, fill_value=(tf[0], tf[-1]))
    k = dist(side) / side
    k = k / k.max() * (1.0 / step)
    return tf_of_k(k)


class SZ_data:
    """Attribute bag of the SZ set-up (reference ``joxsz_funcs.py:136-170``): physical constants
    ``[m_e keV, sigma_T cm^2]``, pixel ``step`` (arcsec), ``kpc_as``, ``convert`` (T keV -> mJy/beam
    per unit y), ``flux_data`` (r, flux, err), ``beam_2d``, ``radius`` (arcsec), ``sep`` (index of
    radius 0), ``r_pp`` (kpc), ``d_mat`` (kpc), ``filtering``, and the optional integrated-Compton prior."""

    def __init__(self, phys_const, step, kpc_as, convert, flux_data, beam_2d, radius, sep, r_pp, d_mat,
                 filtering, calc_integ=False, integ_mu=None, integ_sig=None):
        self.phys_const = phys_const
        self.step = step
        self.kpc_as = kpc_as
        self.convert = convert
        self.flux_data = flux_data
        self.beam_2d = beam_2d
        self.radius = radius
        self.sep = sep
        self.r_pp = r_pp
        self.d_mat = d_mat
        self.filtering = filtering
        self.calc_integ = calc_integ
        self.integ_mu = integ_mu
        self.integ_sig = integ_sig


def getEdges(infg, bands):
    """Annulus edges (arcmin) from the first band's foreground file (reference ``joxsz_funcs.py:172-182``)."""
    tab = np.loadtxt(infg % (bands[0][0], bands[0][1]))
    centre, halfw = tab[:, 0], tab[:, 1]
    return np.concatenate(([centre[0] - halfw[0]], centre + halfw))


def band_from_tables(fg, bg, bandE, rmf, arf):
    """Build a ``Band`` from already-loaded foreground/background tables (see :func:`loadBand`)."""
    centre, halfw, cts, area, expo = (fg[:, i] for i in range(5))
    geom = np.pi * ((centre + halfw) ** 2 - (centre - halfw) ** 2)
    band = mb.Band(bandE[0] / 1000, bandE[1] / 1000, cts, rmf, arf, expo, areascales=area / geom)
    n = centre.size
    band.backrates = bg[:n, 4]
    if abs(bg[:n, 0][-1] - centre[-1]) > 0.001:
        raise RuntimeError("Problem while reading bg file", bg[:n, 0][-1], centre[-1])
    return band


def loadBand(infg, inbg, bandE, rmf, arf):
    """Foreground + background annulus profiles of one energy band -> ``mb.Band``
    (reference ``joxsz_funcs.py:184-211``).  Columns of the foreground file: centre (arcmin),
    half-width, counts, pixelised area (arcmin^2), exposure (s); background: rate in column 5."""
    fg = np.loadtxt(infg % (bandE[0], bandE[1]))
    bg = np.loadtxt(inbg % (bandE[0], bandE[1]))
    return band_from_tables(fg, bg, bandE, rmf, arf)
