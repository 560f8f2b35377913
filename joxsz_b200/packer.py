"""Freeze a built JoXSZ ``fit`` object into the flat constant tables ``jx_create`` consumes.

Input: the object graph ``joxsz_main.py:93-188`` builds -- ``fit.pars`` (name -> Param),
``fit.thawed`` (sampling order, ``joxsz_main.py:179``), ``fit.model`` (ModelNullPot with the Vikhlinin
density / UPP temperature / flat metallicity components), ``fit.data.sz`` (``SZ_data``),
``fit.data.annuli`` and ``fit.data.bands`` (mbproj2 objects or the bundled work-alikes).
Output: a :class:`PackedSetup` whose numpy arrays back the pointers of a ``jx_setup`` struct.

Everything here is one-time host work in float64; nothing evaluates a likelihood.  Geometry the CUDA
kernels do not support raises :class:`operators.GeometryError` here instead of degrading at run time.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib, operators as ops

KPC_CM_DEFAULT = 3.0856776e21


class PackError(ValueError):
    pass


def _kpc_cm():
    try:
        from .mb import mb
        return float(mb.physconstants.kpc_cm)
    except Exception:  # pragma: no cover
        return KPC_CM_DEFAULT


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _convert_table(convert):
    """(T_keV, 1e3*I0) of ``interp1d(t_keV, 1e3*compt_Jy_beam, 'linear', fill_value='extrapolate')``
    (``joxsz_main.py:108-109``)."""
    if isinstance(convert, (tuple, list)) and len(convert) == 2:
        x, y = convert
    else:
        x, y = getattr(convert, "x", None), getattr(convert, "y", None)
        kind = getattr(convert, "_kind", "linear")
        if x is None or y is None or kind != "linear":
            raise PackError("SZ_data.convert must be a linear scipy interp1d (joxsz_main.py:109) or a (T, I) pair")
    x, y = _f64(x), _f64(y)
    if x.ndim != 1 or x.size < 2 or x.shape != y.shape or not np.all(np.diff(x) > 0):
        raise PackError("conversion table must be 1-D, increasing, with >= 2 points")
    return x, y


def integ_weights(r_pp, kpc_as, step):
    """``w`` with ``cint = w @ y`` for the reference's integrated Compton parameter
    ``simps(concatenate((f(0), y)) * x, x) * 2 pi`` (``joxsz_funcs.py:481-483``), ``x`` in arcmin.

    The Simpson weights are taken from the scipy that is installed (``simps`` where it still exists, else
    ``simpson``) by integrating unit vectors, so the even-sample-count rule -- which changed in scipy 1.11 --
    is whatever the user's reference run would apply.  ``x[0] = 0`` removes the ``f(0)`` term."""
    import scipy.integrate as si
    rule = getattr(si, "simps", None) or si.simpson
    r_pp = np.asarray(r_pp, dtype=np.float64)
    x = np.arange(0., r_pp[-1] / kpc_as / 60 + step / 60, step / 60)
    if x.size != r_pp.size + 1:
        raise PackError(f"integration grid has {x.size} samples for {r_pp.size} radii (joxsz_funcs.py:482 would raise)")
    sw = np.asarray(rule(np.eye(x.size), x=x, axis=-1), dtype=np.float64)      # Simpson weight of each sample
    return 2 * np.pi * sw[1:] * x[1:]


class PackedSetup:
    """Numpy tables + the ``jx_setup`` struct that points into them (keep this object alive while
    the struct is in use)."""

    def __init__(self, fit, max_walkers=1024, device=0):
        # The constant operators are products / solves done by BLAS / LAPACK, whose summation order depends on the number
        # of threads the library runs with -- and that differs between a plain `python` process (all cores) and the
        # ranks `torchrun` starts (OMP_NUM_THREADS=1).  Operators that differ in the last bit make the log-likelihoods
        # of the same walker differ in the last bit between a 1-GPU and an N-GPU run (seen on hardware as a
        # `state_checksum` mismatch), which breaks the bit-identity of the chains.  Packing is therefore single-threaded.
        try:
            from threadpoolctl import threadpool_limits
            ctx = threadpool_limits(limits=1)
        except Exception:                      # pragma: no cover - threadpoolctl is a dependency of scipy's wheels
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            self._pack(fit, max_walkers, device)

    def _pack(self, fit, max_walkers, device):
        pars = fit.pars
        thawed = list(fit.thawed)
        sz = fit.data.sz
        annuli = fit.data.annuli
        bands = fit.data.bands
        model = fit.model
        self.thawed = thawed
        self.ndim = len(thawed)
        if self.ndim < 1:
            raise PackError("no thawed parameters")

        # ---------------- parameters
        dens_mode = getattr(model.ne_cmpt, "mode", "single")
        if dens_mode not in ("single", "double"):
            raise PackError(f"unknown density mode {dens_mode!r}")
        self.dens_mode = dens_mode
        z_name = getattr(model.Z_cmpt, "name", "Z")
        slot_names = list(_lib.PARAM_SLOTS)
        slot_names[slot_names.index("Z")] = z_name
        self.slot_names = slot_names
        slot_src = np.full(_lib.JX_NPAR, -1, dtype=np.int32)
        slot_val = np.zeros(_lib.JX_NPAR, dtype=np.float64)
        optional = {"log(n_{02})", r"\beta_2", "log(r_{c2})"} if dens_mode == "single" else set()
        for i, name in enumerate(slot_names):
            if name not in pars:
                if name in optional:
                    continue
                raise PackError(f"fit.pars lacks parameter {name!r}")
            if name in thawed:
                slot_src[i] = thawed.index(name)
            slot_val[i] = float(np.asarray(pars[name].val).reshape(-1)[0])
        self.slot_src, self.slot_val = slot_src, slot_val

        kind = np.zeros(self.ndim, dtype=np.int32)
        pa = np.zeros(self.ndim)
        pb = np.zeros(self.ndim)
        prior_const = 0.0
        for name, par in pars.items():
            gauss = hasattr(par, "prior_mu")
            if name in thawed:
                j = thawed.index(name)
                if gauss:
                    kind[j], pa[j], pb[j] = 1, float(par.prior_mu), float(par.prior_sigma)
                else:
                    kind[j], pa[j], pb[j] = 0, float(par.minval), float(par.maxval)
            else:
                prior_const += float(par.prior())
        if not math.isfinite(prior_const):
            raise PackError("a frozen parameter lies outside its prior: every likelihood would be -inf")
        self.prior_kind, self.prior_a, self.prior_b = kind, pa, pb
        self.prior_const = prior_const

        # ---------------- SZ geometry and operators
        r_pp = _f64(sz.r_pp)
        d_mat = _f64(sz.d_mat)
        self.r_pp = r_pp
        nr = r_pp.size
        mo = ops.SZMapOperators(r_pp, d_mat, sz.beam_2d, sz.filtering, sz.step)
        self.map_ops = mo
        N, H = mo.N, mo.H
        sep = int(sz.sep)
        if sep != H - 1:
            raise ops.GeometryError(f"sep={sep} but the map half-side is {H - 1}: radius and d_mat disagree")
        radius = _f64(sz.radius)
        if radius.size != N:
            raise ops.GeometryError("radius and d_mat have different sides")
        if sep > nr:
            raise ops.GeometryError("sep exceeds len(r_pp)")
        yscale = _kpc_cm() * float(sz.phys_const[1]) / float(sz.phys_const[0])
        self.yscale = yscale
        A = ops.abel_forward_matrix(r_pp)
        self.y_op = _f64(yscale * A)                                   # y = y_op @ pp
        G = ops.sz_spline_coeff_operator(r_pp, mo.nseg)                # coef = G @ y
        self.proj_op = _f64(G @ self.y_op)                             # [4*nseg, nr]
        self.seg = _i32(mo.seg)
        self.dx = _f64(mo.dx)
        self.bhat = _f64(mo.bhat)
        self.bmix = _f64(mo.bmix)
        self.cmat = _f64(mo.cmat)
        self.hf = _f64(mo.hf)
        self.dinv = _f64(mo.dinv)
        self.filt_q = _f64(mo.filt_q)
        # tail
        self.w_t0 = _f64(ops.central_value_operator(r_pp[:sep]))
        self.conv_T, self.conv_I = _convert_table(sz.convert)
        flux = _f64(sz.flux_data)
        if flux.ndim != 2 or flux.shape[0] < 3:
            raise PackError("flux_data must be [>=3, Nd] (radius, flux, error)")
        self.flux_r, self.flux, self.flux_err = (np.ascontiguousarray(flux[i]) for i in range(3))
        self.g_op = _f64(ops.spline_eval_operator(radius[sep:], self.flux_r))     # [Nd, H]
        self.N, self.H, self.sep, self.nr = N, H, sep, nr
        # optional integrated-Compton-parameter penalty (joxsz_funcs.py:480-487)
        self.calc_integ = bool(getattr(sz, "calc_integ", False))
        # the reference only forms the integration grid when calc_integ is on (joxsz_funcs.py:480): a setup whose
        # floating-point arange grid is one sample off must still pack when the penalty is disabled
        if self.calc_integ:
            self.w_integ = _f64(self.y_op.T @ integ_weights(r_pp, float(sz.kpc_as), float(sz.step)))
        else:
            try:
                self.w_integ = _f64(self.y_op.T @ integ_weights(r_pp, float(sz.kpc_as), float(sz.step)))
            except PackError:
                self.w_integ = np.zeros(nr, dtype=np.float64)      # the 'integ' tap then reads 0
        self.integ_mu = float(sz.integ_mu) if getattr(sz, "integ_mu", None) is not None else 0.0
        self.integ_sig = float(sz.integ_sig) if getattr(sz, "integ_sig", None) is not None else 1.0
        if self.calc_integ and not (self.integ_sig > 0 and math.isfinite(self.integ_mu)):
            raise PackError("calc_integ=True needs a finite integ_mu and integ_sig > 0")

        # ---------------- X-ray tables
        na = int(annuli.nshells)
        self.midpt_kpc = _f64(annuli.midpt_kpc)
        self.projvols = _f64(annuli.projvols_cm3)
        geom = _f64(annuli.geomarea_arcmin2)
        ctr = annuli.ctrate
        self.tlog = _f64(ctr.Tlogvals)
        self.tmin, self.tmax = float(ctr.Tmin), float(ctr.Tmax)
        nb = len(bands)
        ntab = self.tlog.size
        ln0 = np.empty((nb, ntab))
        ln1 = np.empty((nb, ntab))
        cts = np.empty((nb, na))
        src = np.empty((nb, na))
        bkg = np.zeros((nb, na))
        nh = model.NH_1022pcm2
        for b, band in enumerate(bands):
            t0, t1 = self._band_tables(ctr, band, nh)
            ln0[b], ln1[b] = t0, t1
            cts[b] = np.asarray(band.cts, dtype=np.float64)
            src[b] = np.asarray(band.areascales, dtype=np.float64) * np.asarray(band.exposures, dtype=np.float64)
            if band.backrates is not None:
                bkg[b] = np.asarray(band.backrates, dtype=np.float64) * geom * src[b]
        self.lnrate0, self.lnrate1 = _f64(ln0), _f64(ln1)
        self.cts, self.srcscale, self.bkgterm = _f64(cts), _f64(src), _f64(bkg)
        self.na, self.nb, self.ntab = na, nb, ntab
        self.exclude_unphy_mass = bool(getattr(fit, "exclude_unphy_mass", False))
        self.max_walkers = int(max_walkers)
        self.device = int(device)
        self._struct = None

    @staticmethod
    def _band_tables(ctr, band, nh):
        if hasattr(ctr, "getTables"):
            return ctr.getTables(band.rmf, band.arf, band.emin_keV, band.emax_keV, nh)
        # real mbproj2: the cache is keyed (emin, emax, z, NH, rmf, arf) -- joxsz_funcs.py:656
        key = (band.emin_keV, band.emax_keV, ctr.cosmo.z, nh, band.rmf, band.arf)
        if key not in ctr.ctcache:
            ctr.addCountCache(key)
        return ctr.ctcache[key]

    # ------------------------------------------------------------------
    def struct(self):
        """The ``jx_setup`` ctypes struct (pointers into this object's arrays)."""
        if self._struct is not None:
            return self._struct
        s = _lib.JxSetup()
        s.abi_version = _lib.JX_ABI_VERSION
        s.device = self.device
        s.max_walkers = self.max_walkers
        s.ndim = self.ndim
        for i in range(_lib.JX_NPAR):
            s.slot_src[i] = int(self.slot_src[i])
            s.slot_val[i] = float(self.slot_val[i])
        s.dens_mode = 1 if self.dens_mode == "double" else 0
        s.exclude_unphy_mass = int(self.exclude_unphy_mass)
        s.prior_const = self.prior_const
        mo = self.map_ops
        s.nr, s.nt, s.nmap, s.nh, s.npad, s.nseg = self.nr, self.sep, self.N, self.H, mo.P, mo.nseg
        s.nconv, s.nd = self.conv_T.size, self.flux.size
        s.na, s.nb, s.ntab = self.na, self.nb, self.ntab
        s.tmin, s.tmax = self.tmin, self.tmax

        def pd(a):
            assert a.dtype == np.float64 and a.flags.c_contiguous
            return a.ctypes.data_as(C.POINTER(C.c_double))

        def pi(a):
            assert a.dtype == np.int32 and a.flags.c_contiguous
            return a.ctypes.data_as(C.POINTER(C.c_int32))

        s.prior_kind, s.prior_a, s.prior_b = pi(self.prior_kind), pd(self.prior_a), pd(self.prior_b)
        s.r_pp, s.proj_op, s.y_op = pd(self.r_pp), pd(self.proj_op), pd(self.y_op)
        s.seg, s.dx, s.bhat = pi(self.seg), pd(self.dx), pd(self.bhat)
        s.nbeam, s.bmix = int(self.bmix.shape[0]), pd(self.bmix)
        s.cmat, s.hf, s.dinv, s.filt_q = pd(self.cmat), pd(self.hf), pd(self.dinv), pd(self.filt_q)
        s.w_t0, s.conv_T, s.conv_I = pd(self.w_t0), pd(self.conv_T), pd(self.conv_I)
        s.g_op, s.flux, s.flux_err = pd(self.g_op), pd(self.flux), pd(self.flux_err)
        s.calc_integ, s.w_integ = int(self.calc_integ), pd(self.w_integ)
        s.integ_mu, s.integ_sig = self.integ_mu, self.integ_sig
        s.midpt_kpc, s.projvols, s.tlog = pd(self.midpt_kpc), pd(self.projvols), pd(self.tlog)
        s.lnrate0, s.lnrate1 = pd(self.lnrate0), pd(self.lnrate1)
        s.cts, s.srcscale, s.bkgterm = pd(self.cts), pd(self.srcscale), pd(self.bkgterm)
        self._struct = s
        return s

    # algorithmic per-walker figures used by bench.py
    def algorithmic_bytes(self):
        s = 8
        N, nr, H = self.N, self.nr, self.H
        return {
            "profiles": s * (self.ndim + nr + H + 2 * self.na + 1),
            "project": s * (nr + 4 * self.map_ops.nseg),
            "szmap": s * (nr + 3 * N * N + H),
            "xray": s * (3 * self.na + self.nb * self.na),
            "tail": s * (H + self.nb * self.na + 1),
        }

    def algorithmic_flops(self):
        N, nr = self.N, self.nr
        pd = self.map_ops.P
        lg = math.log2
        return {
            "project": 2.0 * nr * 4 * self.map_ops.nseg,
            # filter stage: one GEMM row of K = H (H + 1) / 2 distinct convolved-map pixels by H outputs (k7_filter.cu)
            "filter": 2.0 * (self.H * (self.H + 1) // 2) * self.H,
            # SURVEY 8(d) convention (full complex FFTs): padded convolution + exact-size filter.  The map kernels end
            # at the convolved map, so only the first half is their own
            "szmap": 2 * 5 * pd * pd * lg(pd * pd) + 6 * pd * pd,
            "szmap_plus_filter": 2 * 5 * pd * pd * lg(pd * pd) + 2 * 5 * N * N * lg(N * N) + 6 * pd * pd + 6 * N * N,
        }
