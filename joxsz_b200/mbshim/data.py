"""Geometry and data containers (``mbproj2.Annuli/Band/Data`` work-alikes)."""
import numpy as np

from . import utils
from .countrate import CountRate
from .physconstants import kpc_cm


class Annuli:
    def __init__(self, edges_arcmin, cosmology):
        self.cosmology = cosmology
        self.edges_arcmin = np.asarray(edges_arcmin, dtype=np.float64)
        self.nshells = len(self.edges_arcmin) - 1
        e = self.edges_arcmin
        self.geomarea_arcmin2 = np.pi * (e[1:] ** 2 - e[:-1] ** 2)
        self.update()

    def update(self):
        e_cm = self.cosmology.kpc_per_arcsec * self.edges_arcmin * 60.0 * kpc_cm
        self.edges_cm = e_cm
        self.rin_cm = rin = e_cm[:-1]
        self.rout_cm = rout = e_cm[1:]
        self.midpt_cm = 0.5 * (rin + rout)
        self.massav_cm = 0.75 * (rout ** 4 - rin ** 4) / (rout ** 3 - rin ** 3)
        self.widths_cm = rout - rin
        self.vols_cm3 = 4.0 / 3.0 * np.pi * (rout ** 3 - rin ** 3)
        self.projvols_cm3 = np.ascontiguousarray(utils.projectionVolumeMatrix(e_cm).T)
        for nm in ("edges", "rin", "rout", "midpt", "massav", "widths"):
            v = getattr(self, nm + "_cm") / kpc_cm
            setattr(self, nm + "_kpc", v)
            with np.errstate(divide="ignore"):
                setattr(self, nm + "_logkpc", np.log(v))
        self.ctrate = CountRate(self.cosmology)


class Band:
    def __init__(self, emin_keV, emax_keV, cts, rmf, arf, exposures, backrates=None, areascales=None):
        self.emin_keV = emin_keV
        self.emax_keV = emax_keV
        self.cts = np.asarray(cts, dtype=np.float64)
        self.rmf = rmf
        self.arf = arf
        self.exposures = np.asarray(exposures, dtype=np.float64)
        self.backrates = backrates
        self.areascales = np.ones_like(self.cts) if areascales is None else np.asarray(areascales, dtype=np.float64)

    def calcProjProfile(self, annuli, ne_prof, T_prof, Z_prof, NH_1022pcm2, backscale=1.0):
        """mbproj2 projects the shell count rates through ``projvols_cm3`` here on the host.  In this package the
        projection exists only in the CUDA kernel K4 (``jx_xray``): use ``fit.calcProfiles()``."""
        raise NotImplementedError("Band.calcProjProfile has no host implementation here: predicted profiles are "
                                  "computed on the GPU (jx_xray); call fit.calcProfiles()")


class Data:
    def __init__(self, bands, annuli):
        for b in bands:
            assert len(b.cts) == annuli.nshells
        self.bands = bands
        self.annuli = annuli
