"""Import the UNMODIFIED reference ``/root/reference/joxsz_funcs.py`` in this container.

Its third-party imports that are not installable here are stubbed in ``sys.modules``:

* ``astropy.io.fits``  -> a two-line adapter over ``joxsz_b200.fitsio`` (same ``open(f)[''].data[0]`` shape)
* ``mbproj2``          -> ``joxsz_b200.mbshim`` (restated from memory)
* ``abel.direct``      -> ``oracle.joxsz_oracle.pyabel_direct_forward`` (restated; Appendix A.1)
* ``h5py``             -> empty module (only used by add_backend_attrs / addCountCache)
* ``scipy.integrate.simps`` -> ``simpson`` (removed from scipy >= 1.14; only used when calc_integ=True)

Consequence, stated wherever the golden vectors are used: they pin this repo's oracle against the
reference's OWN code, not against PyAbel / mbproj2 themselves.
"""
import importlib.util
import os
import sys
import types

REFERENCE_DIR = "/root/reference"


def reference_available():
    return os.path.exists(os.path.join(REFERENCE_DIR, "joxsz_funcs.py"))


def real_deps_available():
    """Which of the third-party packages the reference imports are really installed (never true in the build
    container; a maintainer's machine may have them)."""
    import importlib.util
    out = {}
    for name in ("abel", "mbproj2", "astropy", "h5py"):
        try:
            out[name] = importlib.util.find_spec(name) is not None
        except (ImportError, ValueError):
            out[name] = False
    return out


#: filled by _install_stubs: {"abel": "real" | "stub", "mbproj2": ..., "astropy": ..., "h5py": ...}
MODE = {}


def _install_stubs(real_deps=False):
    """``real_deps=True``: every package of ``real_deps_available()`` that is installed is used as it is (PyAbel's own
    ``direct_transform``, mbproj2's own ``Fit`` / ``Band`` / ``CountRate`` ...) and only the missing ones are stubbed;
    the goldens generated that way pin the third-party boundary itself (make_golden_reference.py --real-deps)."""
    have = real_deps_available() if real_deps else {}
    MODE.clear()
    MODE.update({k: ("real" if have.get(k) else "stub") for k in ("abel", "mbproj2", "astropy", "h5py")})
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.abspath(os.path.join(here, "..", ".."))
    if root not in sys.path:
        sys.path.insert(0, root)
    from joxsz_b200 import fitsio, mbshim
    from oracle import joxsz_oracle as orc

    class _HDU:
        def __init__(self, rows):
            self.data = rows

    class _HDUList:
        def __init__(self, filename):
            self._rows = fitsio.read_bintable(filename, ext=1)

        def __getitem__(self, key):
            return _HDU(self._rows)

    if MODE["astropy"] == "stub":
        fits = types.ModuleType("astropy.io.fits")
        fits.open = lambda filename: _HDUList(filename)
        astropy = types.ModuleType("astropy")
        astropy_io = types.ModuleType("astropy.io")
        astropy.io = astropy_io
        astropy_io.fits = fits
        sys.modules.update({"astropy": astropy, "astropy.io": astropy_io, "astropy.io.fits": fits})

    if MODE["mbproj2"] == "real":
        import mbproj2 as real_mb
        mb_out = real_mb
    else:
        mb_out = mbshim
        _patch_shim_host_arithmetic(mbshim)
        sys.modules["mbproj2"] = mbshim
        sys.modules["mbproj2.physconstants"] = mbshim.physconstants

    if MODE["abel"] == "stub":
        abel = types.ModuleType("abel")
        direct = types.ModuleType("abel.direct")

        def direct_transform(fr, dr=None, r=None, direction="inverse", derivative=None, int_func=None,
                             correction=True, backend="C", **kw):
            assert direction == "forward" and backend == "Python" and r is not None
            return orc.pyabel_direct_forward(fr, r)

        direct.direct_transform = direct_transform
        abel.direct = direct
        sys.modules.update({"abel": abel, "abel.direct": direct})
    if MODE["h5py"] == "stub":
        sys.modules["h5py"] = types.ModuleType("h5py")

    import scipy.integrate
    if not hasattr(scipy.integrate, "simps"):
        scipy.integrate.simps = scipy.integrate.simpson
    return mb_out


def _patch_shim_host_arithmetic(mbshim):
    # The product shim deliberately has no host arithmetic for the per-walker X-ray path; the reference's own
    # getLikelihood needs it, so numpy restatements of the three mbproj2 routines
    # are patched in for this run only.
    import numpy as np

    def _getCountRate(self, rmf, arf, minenergy_keV, maxenergy_keV, NH_1022, T_keV, Z_solar, ne_cm3):
        t0, t1 = self.getTables(rmf, arf, minenergy_keV, maxenergy_keV, NH_1022)
        logT = np.log(np.clip(T_keV, self.Tmin, self.Tmax))
        r0 = np.exp(np.interp(logT, self.Tlogvals, t0))
        r1 = np.exp(np.interp(logT, self.Tlogvals, t1))
        return (r0 + (r1 - r0) * Z_solar) * ne_cm3 ** 2

    def _calcProjProfile(self, annuli, ne_prof, T_prof, Z_prof, NH_1022pcm2, backscale=1.0):
        rates = annuli.ctrate.getCountRate(self.rmf, self.arf, self.emin_keV, self.emax_keV,
                                           NH_1022pcm2, T_prof, Z_prof, ne_prof)
        projrates = annuli.projvols_cm3.dot(rates)
        projrates = projrates * (self.areascales * self.exposures)
        if self.backrates is not None:
            projrates = projrates + (self.backrates * backscale * annuli.geomarea_arcmin2
                                     * self.areascales * self.exposures)
        return projrates

    def _cashLogLikelihood(data, model):
        like = np.sum(data * np.log(model)) - np.sum(model)
        return like if np.isfinite(like) else -np.inf

    def _calcProfiles(self):
        ne_prof, T_prof, Z_prof = self.model.computeProfs(self.pars)
        return [band.calcProjProfile(self.data.annuli, ne_prof, T_prof, Z_prof,
                                     self.model.NH_1022pcm2, backscale=self.pars["backscale"].val)
                for band in self.data.bands]

    mbshim.countrate.CountRate.getCountRate = _getCountRate
    mbshim.Band.calcProjProfile = _calcProjProfile
    mbshim.utils.cashLogLikelihood = _cashLogLikelihood
    mbshim.Fit.calcProfiles = _calcProfiles



def import_reference(real_deps=False):
    """Returns (ref_module, mb) with the reference's joxsz_funcs loaded from /root/reference."""
    mb = _install_stubs(real_deps)
    spec = importlib.util.spec_from_file_location("ref_joxsz_funcs", os.path.join(REFERENCE_DIR, "joxsz_funcs.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref, mb
