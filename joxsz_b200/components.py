"""JoXSZ profile components with the reference's class/method names and signatures.

Reference: ``joxsz_funcs.py:213-437`` (``add_param_unit``, ``Z_defPars``, ``CmptPressure``,
``CmptUPPTemperature``, ``mydens_defPars/vikhFunction/prior``, ``CmptMyMass``).  Parameter names,
defaults, bounds and units are the reference's.  The numerical methods (``press_fun``,
``press_derivative``, ``temp_fun``, ``vikhFunction``, ``mass_fun``) evaluate on the GPU through the
C-ABI entry ``jx_radial_profiles`` (see ``include/joxsz_b200.h``): ``pars[name].val`` may be a python
float or an array of W values (one per walker) and the result is ``[n_r]`` or ``[W, n_r]``.
There is no host implementation of the formulas in this package.
"""
from __future__ import annotations

import numpy as np

from . import radial
from .mb import mb

# ---- parameter names (identical strings to the reference; they are the keys of fit.pars)
P0, A_, B_, C_, RP = "P_0", "a", "b", "c", "r_p"
LN0, BETA, LRC, LRS = "log(n_0)", r"\beta", "log(r_c)", "log(r_s)"
ALPHA, EPS, GAMMA = r"\alpha", r"\epsilon", r"\gamma"
LN02, BETA2, LRC2 = "log(n_{02})", r"\beta_2", "log(r_{c2})"
LTR, ZMET, BACK, CAL = "log(T_X/T_{SZ})", "Z", "backscale", "calibration"


def add_param_unit():
    """Give ``mb.Param`` / ``mb.ParamGaussian`` a ``unit`` attribute (reference ``joxsz_funcs.py:213-239``);
    the Gaussian also gains ignored ``minval``/``maxval`` attributes.  ``prior()`` is untouched."""
    def _param_init(self, val, minval=-1e99, maxval=1e99, unit=".", frozen=False):
        mb.ParamBase.__init__(self, val, frozen=frozen)
        self.minval, self.maxval, self.unit = minval, maxval, unit

    def _param_repr(self):
        return "<Param: val=%.3g, minval=%.3g, maxval=%.3g, unit=%s, frozen=%s>" % (
            self.val, self.minval, self.maxval, self.unit, self.frozen)

    def _gauss_init(self, val, prior_mu, prior_sigma, unit=".", frozen=False, minval=None, maxval=None):
        mb.ParamBase.__init__(self, val, frozen=frozen)
        self.prior_mu, self.prior_sigma, self.unit = prior_mu, prior_sigma, unit
        self.minval, self.maxval = minval, maxval

    def _gauss_repr(self):
        return "<ParamGaussian: val=%.3g, prior_mu=%.3g, prior_sigma=%.3g, unit=%s, frozen=%s>" % (
            self.val, self.prior_mu, self.prior_sigma, self.unit, self.frozen)

    mb.Param.__init__, mb.Param.__repr__ = _param_init, _param_repr
    mb.ParamGaussian.__init__, mb.ParamGaussian.__repr__ = _gauss_init, _gauss_repr


def Z_defPars(self):
    """Metallicity default with a unit (reference ``joxsz_funcs.py:241-246``)."""
    return {self.name: mb.Param(self.defval, unit="solar", minval=self.minval, maxval=self.maxval)}


def _density_mode(ne_prof):
    return getattr(ne_prof, "mode", "single")


class CmptPressure(mb.Cmpt):
    """gNFW pressure profile (reference ``joxsz_funcs.py:248-301``)."""

    def __init__(self, name, annuli):
        mb.Cmpt.__init__(self, name, annuli)

    def defPars(self):
        return {
            P0: mb.Param(0.4, minval=0.0, maxval=2.0, unit="keV.cm^{-3}"),
            A_: mb.Param(1.33, minval=0.1, maxval=20.0, unit="."),
            B_: mb.Param(4.13, minval=0.1, maxval=15.0, unit="."),
            C_: mb.Param(0.014, minval=0.0, maxval=3.0, unit="."),
            RP: mb.Param(300.0, minval=100.0, maxval=3000.0, unit="kpc"),
        }

    def press_fun(self, pars, r_kpc):
        return radial.evaluate(pars, r_kpc, "press", need=(P0, A_, B_, C_, RP))

    def press_derivative(self, pars, r_kpc):
        return radial.evaluate(pars, r_kpc, "dpress", need=(P0, A_, B_, C_, RP))


def mydens_defPars(self):
    """Vikhlinin density defaults (reference ``joxsz_funcs.py:341-373``)."""
    pars = {
        LN0: mb.Param(-3.0, minval=-7.0, maxval=2.0, unit="log(cm^{-3})"),
        BETA: mb.Param(2 / 3, minval=0.0, maxval=4.0, unit="."),
        LRC: mb.Param(2.3, minval=-1.0, maxval=3.7, unit="log(kpc)"),
        LRS: mb.Param(2.7, minval=0.0, maxval=3.7, unit="log(kpc)"),
        ALPHA: mb.Param(0.0, minval=-1.0, maxval=2.0, unit="."),
        EPS: mb.Param(3.0, minval=0.0, maxval=5.0, unit="."),
        GAMMA: mb.Param(3.0, minval=0.0, maxval=10.0, frozen=True, unit="."),
    }
    if self.mode == "double":
        pars.update({
            LN02: mb.Param(-1.0, minval=-7.0, maxval=2.0, unit="log(cm^{-3})"),
            BETA2: mb.Param(0.5, minval=0.0, maxval=4.0, unit="."),
            LRC2: mb.Param(1.7, minval=-1.0, maxval=3.7, unit="log(kpc)"),
        })
    return pars


_DENS_SINGLE = (LN0, BETA, LRC, LRS, ALPHA, EPS, GAMMA)
_DENS_DOUBLE = _DENS_SINGLE + (LN02, BETA2, LRC2)


def _dens_need(mode):
    return _DENS_DOUBLE if mode == "double" else _DENS_SINGLE


def mydens_vikhFunction(self, pars, radii_kpc):
    """Vikhlinin electron density (reference ``joxsz_funcs.py:375-395``)."""
    return radial.evaluate(pars, radii_kpc, "ne", need=_dens_need(self.mode), mode=self.mode)


def mydens_prior(self, pars):
    """-inf when the core radius exceeds the scale radius (reference ``joxsz_funcs.py:397-407``)."""
    rc = np.asarray(pars[LRC].val, dtype=np.float64)
    rs = np.asarray(pars[LRS].val, dtype=np.float64)
    out = np.where(10 ** rc > 10 ** rs, -np.inf, 0.0)
    return float(out) if out.ndim == 0 else out


class CmptUPPTemperature(mb.Cmpt):
    """T_SZ = P / n_e and T_X = T_SZ * 10**log(T_X/T_SZ) (reference ``joxsz_funcs.py:303-339``)."""

    def __init__(self, name, annuli, press_prof, ne_prof):
        mb.Cmpt.__init__(self, name, annuli)
        self.press_prof = press_prof
        self.ne_prof = ne_prof

    def defPars(self):
        return {LTR: mb.Param(0.0, minval=-1.0, maxval=1.0, unit=".")}

    def temp_fun(self, pars, r_kpc, getT_SZ=False):
        mode = _density_mode(self.ne_prof)
        need = (P0, A_, B_, C_, RP) + _dens_need(mode)
        if getT_SZ:
            return radial.evaluate(pars, r_kpc, "tsz", need=need, mode=mode)
        return radial.evaluate(pars, r_kpc, "tx", need=need + (LTR,), mode=mode)

    def computeProf(self, pars):
        return self.temp_fun(pars, self.annuli.midpt_kpc)


class CmptMyMass(mb.Cmpt):
    """Hydrostatic-equilibrium mass profile (reference ``joxsz_funcs.py:409-437``)."""

    def __init__(self, name, annuli, press_prof, ne_prof):
        mb.Cmpt.__init__(self, name, annuli)
        self.press_prof = press_prof
        self.ne_prof = ne_prof

    def defPars(self):
        pars = self.press_prof.defPars()
        pars.update(self.ne_prof.defPars())
        return pars

    def mass_fun(self, pars, r_kpc, mu_gas=0.61):
        mode = _density_mode(self.ne_prof)
        need = (P0, A_, B_, C_, RP) + _dens_need(mode)
        return radial.evaluate(pars, r_kpc, "mass", need=need, mode=mode, mu_gas=mu_gas)
