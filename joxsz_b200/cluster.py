"""Build the JoXSZ ``fit`` object for a cluster, following the set-up part of ``joxsz_main.main()``
(reference ``joxsz_main.py:93-188``) with this package's readers/components.

``ClusterInputs`` holds the raw decoded contents of the data files (or synthetic stand-ins), so the
same builder serves: the shipped CL J1226.9+3332 example read from a JoXSZ data directory, the small
committed fixture ``tests/golden/cl1226_inputs.npz`` (what travels to the GPU box), and synthetic
clusters of other sizes.
"""
from __future__ import annotations

import os
from types import MethodType

import numpy as np
from scipy.interpolate import interp1d

from . import components as cmp
from . import readers
from .mb import mb
from .synthetic import synthetic_countrate_tables

# configuration globals of joxsz_main.py:21-88, kept under the same names
DEFAULTS = dict(
    mystep=2.0, m_e=0.5109989 * 1e3, sigma_T=6.6524587158 * 1e-25, R_b=5000.0,
    redshift=0.888, H0=67.32, WM=0.3158, WV=0.6842,
    bandEs=[[700, 1000], [1000, 1300], [1300, 1600], [1600, 2000], [2000, 2700],
            [2700, 3400], [3400, 3800], [3800, 4300], [4300, 5000], [5000, 7000]],
    NH_1022pcm2=0.0183, Z_solar=0.3, exclude_unphy_mass=True,
    calc_integ=False, integ_mu=.94 / 1e3, integ_sig=.36 / 1e3,
    rmf="source.rmf", arf="source.arf",
)


class ClusterInputs:
    """Raw inputs: ``beam_r``/``beam_b`` (tabulated beam, or None with ``beam_fwhm`` for a Gaussian),
    ``tf_wn``/``tf`` (transfer function), ``flux_data`` [3, Nd], ``conv_T``/``conv_I`` (keV, Jy/beam per y),
    ``xfg``/``xbg`` [nb, na, 5] (foreground / background annulus tables), plus the config globals."""

    def __init__(self, **kw):
        cfg = dict(DEFAULTS)
        cfg.update(kw)
        self.__dict__.update(cfg)


def load_cl1226_files(data_dir, **overrides):
    """Read the shipped example from a JoXSZ ``data/`` directory (``joxsz_main.py:52-56, 81-85``)."""
    szd, xd = os.path.join(data_dir, "SZ"), os.path.join(data_dir, "X")
    beam_r, beam_b = readers.read_beam(os.path.join(szd, "Beam150GHz.fits"))
    tf_wn, tf = readers.read_tf(os.path.join(szd, "TransferFunction150GHz_CLJ1227.fits"))
    flux = np.asarray(readers.read_xy_err(os.path.join(szd, "press_data_cl1226_flagsource_Xraycent.dat"), ncol=3))
    conv_T, conv_I = np.loadtxt(os.path.join(szd, "Compton_to_Jy_per_beam.dat"), skiprows=1, unpack=True)
    bandEs = overrides.get("bandEs", DEFAULTS["bandEs"])
    xfg = np.stack([np.loadtxt(os.path.join(xd, "fg_profnew_%04i_%04i.dat" % tuple(b))) for b in bandEs])
    xbg = np.stack([np.loadtxt(os.path.join(xd, "bg_profnew_%04i_%04i.dat" % tuple(b))) for b in bandEs])
    return ClusterInputs(beam_r=beam_r, beam_b=beam_b, beam_fwhm=None, tf_wn=tf_wn, tf=tf, flux_data=flux,
                         conv_T=conv_T, conv_I=conv_I, xfg=xfg, xbg=xbg, **overrides)


_NPZ_KEYS = ("beam_r", "beam_b", "tf_wn", "tf", "flux_data", "conv_T", "conv_I", "xfg", "xbg")


def save_inputs_npz(inp, path):
    np.savez_compressed(path, **{k: np.asarray(getattr(inp, k)) for k in _NPZ_KEYS})


def load_inputs_npz(path, **overrides):
    z = np.load(path)
    return ClusterInputs(beam_fwhm=None, **{k: z[k] for k in _NPZ_KEYS}, **overrides)


def synthetic_inputs(map_half=128, nr=512, base=None, n_sz=None, seed=7, **overrides):
    """A synthetic cluster of another size: pixel step 2", exactly 8 kpc/", ``r_pp = h*[1..nr]``,
    map side ``2*map_half+1``, Gaussian beam of FWHM 18.5" (``mybeam(approx=True)``), transfer function
    ``0.95*Phi(k/0.02)`` (``read_tf(approx=True)``, joxsz_funcs.py:100-101), SZ points every ~3 pixels with
    the shipped file's error pattern, and the shipped X-ray layout (``base`` inputs) if given."""
    rng = np.random.default_rng(seed)
    step = overrides.get("mystep", DEFAULTS["mystep"])
    fwhm = 18.5
    maxr_data = map_half * step - 3 * fwhm          # so that (maxr_data + 3 fwhm)//step*step = map_half*step
    n_sz = n_sz or max(8, (2 * map_half + 1) // 9)
    r_sz = np.linspace(1.5 * step, maxr_data, n_sz)
    r_sz[-1] = maxr_data
    err = 0.08 + 0.12 * (r_sz / r_sz[-1]) + 0.02 * rng.random(n_sz)
    prof = -2.4 * (1 + (r_sz / 40.0) ** 2) ** -0.9
    flux = np.stack([r_sz, prof + err * rng.standard_normal(n_sz), err])
    wn = np.linspace(0.0, 0.5, 76)
    kw = dict(beam_r=None, beam_b=None, beam_fwhm=fwhm, tf_wn=wn, tf=None, tf_approx=(0.0, 0.02, 0.95),
              flux_data=flux, kpc_as=8.0, r_pp_count=nr)
    if base is not None:
        kw.update(conv_T=base.conv_T, conv_I=base.conv_I, xfg=base.xfg, xbg=base.xbg)
    else:
        kw.update(conv_T=np.array([1., 5., 10., 15., 20., 25.]),
                  conv_I=np.array([-11.63, -11.34, -11.00, -10.71, -10.38, -10.17]),
                  xfg=_synthetic_xray_fg(rng), xbg=None)
        kw["xbg"] = _synthetic_xray_bg(kw["xfg"])
    kw.update(overrides)
    return ClusterInputs(**kw)


def _synthetic_xray_fg(rng, nb=10):
    edges = np.array([0, .05, .1, .15, .2, .25, .3, .4, .5, 1, 1.3333, 2, 2.6667, 4.3333, 6, 7.6667])
    centre, halfw = 0.5 * (edges[1:] + edges[:-1]), 0.5 * (edges[1:] - edges[:-1])
    area = np.pi * (edges[1:] ** 2 - edges[:-1] ** 2) * 0.9
    fg = np.zeros((nb, centre.size, 5))
    for b in range(nb):
        fg[b, :, 0], fg[b, :, 1] = centre, halfw
        fg[b, :, 2] = rng.poisson(40.0 / (1 + b) + 5, size=centre.size)
        fg[b, :, 3], fg[b, :, 4] = area, 2.5e4
    return fg


def _synthetic_xray_bg(fg):
    bg = np.zeros_like(fg)
    bg[:, :, 0] = fg[:, :, 0]
    bg[:, :, 4] = 1.3e-4
    return bg


def build_fit(inp: ClusterInputs, tables="synthetic", savedir="./"):
    """Return ``(fit, sz_data)`` built exactly in the order of ``joxsz_main.py:93-188``.

    ``tables``: "synthetic" fills the count-rate cache with :func:`synthetic_countrate_tables`
    (XSPEC is unavailable); a list of ``(lnrate_Z0, lnrate_Z1)`` per band uses those; None leaves the
    cache empty (real mbproj2 + XSPEC would build it on first use).
    """
    cosmology = mb.Cosmology(inp.redshift)
    cosmology.H0, cosmology.WM, cosmology.WV = inp.H0, inp.WM, inp.WV
    mystep = inp.mystep
    phys_const = [inp.m_e, inp.sigma_T]
    kpc_as = getattr(inp, "kpc_as", None) or cosmology.kpc_per_arcsec
    flux_data = np.asarray(inp.flux_data, dtype=np.float64)
    maxr_data = flux_data[0][-1]
    if inp.beam_r is not None:
        beam_2d, fwhm = readers.beam_image(mystep, maxr_data, inp.beam_r, inp.beam_b)
    else:
        beam_2d, fwhm = readers.beam_image(mystep, maxr_data, approx=True, fwhm_beam=inp.beam_fwhm)
    mymaxr = (maxr_data + 3 * fwhm) // mystep * mystep
    radius = np.arange(0.0, mymaxr + mystep, mystep)
    radius = np.append(-radius[:0:-1], radius)
    sep = radius.size // 2
    h = mystep * kpc_as
    nr_fixed = getattr(inp, "r_pp_count", None)
    r_pp = h * np.arange(1, nr_fixed + 1) if nr_fixed else np.arange(h, inp.R_b + h, h)
    d_mat = readers.centdistmat(radius * kpc_as)
    tf = inp.tf
    if tf is None:
        from scipy.stats import norm
        loc, scale, c = inp.tf_approx
        tf = c * norm.cdf(inp.tf_wn, loc, scale)
    filtering = readers.filt_image(inp.tf_wn, tf, d_mat.shape[0], mystep)
    convert = interp1d(inp.conv_T, 1e3 * np.asarray(inp.conv_I), "linear", fill_value="extrapolate")
    sz_data = readers.SZ_data(phys_const, mystep, kpc_as, convert, flux_data, beam_2d, radius, sep, r_pp, d_mat,
                              filtering, inp.calc_integ, inp.integ_mu, inp.integ_sig)

    xfg, xbg = np.asarray(inp.xfg), np.asarray(inp.xbg)
    edges = np.concatenate(([xfg[0, 0, 0] - xfg[0, 0, 1]], xfg[0, :, 0] + xfg[0, :, 1]))
    annuli = mb.Annuli(edges, cosmology)
    bands = [readers.band_from_tables(xfg[i], xbg[i], bandE, inp.rmf, inp.arf) for i, bandE in enumerate(inp.bandEs)]
    data = mb.Data(bands, annuli)
    data.sz = sz_data

    cmp.add_param_unit()
    Z_cmpt = mb.CmptFlat("Z", annuli, defval=inp.Z_solar, minval=0.0, maxval=1.0)
    mb.CmptFlat.defPars = cmp.Z_defPars
    ne_cmpt = mb.CmptVikhDensity("ne", annuli, mode=getattr(inp, "dens_mode", "single"))
    mb.CmptVikhDensity.vikhFunction = cmp.mydens_vikhFunction
    mb.CmptVikhDensity.defPars = cmp.mydens_defPars
    mb.CmptVikhDensity.prior = cmp.mydens_prior
    press_cmpt = cmp.CmptPressure("p", annuli)
    T_cmpt = cmp.CmptUPPTemperature("T", annuli, press_cmpt, ne_cmpt)
    model = mb.ModelNullPot(annuli, ne_cmpt, T_cmpt, Z_cmpt, NH_1022pcm2=inp.NH_1022pcm2)
    pars = model.defPars()
    pars.update(press_cmpt.defPars())
    pars["backscale"] = mb.ParamGaussian(1.0, prior_mu=1.0, prior_sigma=0.1)
    pars["calibration"] = mb.ParamGaussian(1.0, prior_mu=1.0, prior_sigma=0.07)
    pars["log(r_c)"].maxval = annuli.edges_logkpc[-2]
    pars["log(r_s)"].maxval = annuli.edges_logkpc[-2]
    pars[r"\gamma"].val = 3.0
    pars[r"\gamma"].frozen = True
    pars["log(r_c)"].val = 2.0
    pars[r"\epsilon"].maxval = 10.0
    pars[r"\alpha"].val = 0.0
    pars[r"\alpha"].frozen = True
    pars["c"].frozen = True
    pars["log(T_X/T_{SZ})"].frozen = False

    fit = mb.Fit(pars, model, data)
    fit.thawed = [name for name, par in fit.pars.items() if not par.frozen]
    fit.exclude_unphy_mass = inp.exclude_unphy_mass
    fit.savedir = savedir
    fit.press = press_cmpt
    fit.mass_cmpt = cmp.CmptMyMass("m", annuli, press_cmpt, ne_cmpt)

    if tables is not None:
        ctr = annuli.ctrate
        if isinstance(tables, str) and tables == "synthetic":
            tables = synthetic_countrate_tables([(b.emin_keV, b.emax_keV) for b in bands], ctr.Tlogvals)
        for band, (t0, t1) in zip(bands, tables):
            key = (band.emin_keV, band.emax_keV, cosmology.z, inp.NH_1022pcm2, band.rmf, band.arf)
            ctr.ctcache[key] = (np.asarray(t0, dtype=np.float64), np.asarray(t1, dtype=np.float64))

    from . import funcs
    # joxsz_main.py:186-188 binds these on the CLASS (so they always act on the last fit built); binding on
    # the instance gives the same calls for a single fit and keeps several fits in one process independent
    fit.get_sz_like = MethodType(funcs.get_sz_like, fit)
    fit.getLikelihood = MethodType(funcs.getLikelihood, fit)
    fit.mylikeFromProfs = MethodType(funcs.mylikeFromProfs, fit)
    fit.calcProfiles = MethodType(funcs.calcProfiles, fit)
    from . import fitting
    fit.doFitting = MethodType(fitting.doFitting, fit)      # batched multi-start simplex instead of the serial scipy loop
    return fit, sz_data
