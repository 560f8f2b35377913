"""``mbproj2.countrate.CountRate`` work-alike: tabulated count rate vs. ln T at Z=0 and Z=1.

In mbproj2 the two tables per band are built by XSPEC (``phabs*apec``, ne=1) and cached; the
format is documented by the reference's own ``addCountCache`` (``joxsz_funcs.py:652-681``):
``(ln rate_Z0[Tsteps], ln rate_Z1[Tsteps])`` on the natural-log grid ``Tlogvals``.  XSPEC cannot
run here, so tables are supplied explicitly with :meth:`setTables` (synthetic ones for the tests
and the bench, see ``joxsz_b200.synthetic``).  Evaluation rule::

    rate = (exp(interp(lnT, Tlog, t0)) + (exp(interp(lnT, Tlog, t1)) - exp(interp(..t0))) * Z) * ne**2
"""
import numpy as np


class CountRate:
    Tmin = 0.06
    Tmax = 60.0
    Tsteps = 100
    Tlogvals = np.linspace(np.log(Tmin), np.log(Tmax), Tsteps)

    def __init__(self, cosmo):
        self.cosmo = cosmo
        self.ctcache = {}

    @staticmethod
    def makeKey(rmf, arf, minenergy_keV, maxenergy_keV, NH_1022, z):
        return (minenergy_keV, maxenergy_keV, z, NH_1022, rmf, arf)

    def setTables(self, key, lnrate_Z0, lnrate_Z1):
        self.ctcache[key] = (np.asarray(lnrate_Z0, dtype=np.float64),
                             np.asarray(lnrate_Z1, dtype=np.float64))

    def addCountCache(self, key):
        raise RuntimeError(
            "count-rate table for %r is missing and XSPEC is not available; call setTables()" % (key,))

    def getTables(self, rmf, arf, minenergy_keV, maxenergy_keV, NH_1022):
        key = self.makeKey(rmf, arf, minenergy_keV, maxenergy_keV, NH_1022, self.cosmo.z)
        if key not in self.ctcache:
            self.addCountCache(key)
        return self.ctcache[key]

    def getCountRate(self, rmf, arf, minenergy_keV, maxenergy_keV, NH_1022, T_keV, Z_solar, ne_cm3):
        """mbproj2 evaluates ``(exp(interp0) + (exp(interp1) - exp(interp0)) Z) ne^2`` here on the host.  In this
        package that arithmetic exists only in the CUDA kernel K4 (``jx_xray``): use ``fit.calcProfiles()`` /
        ``BatchedLikelihood.xray``.  (The golden-vector generator patches a numpy restatement in for the run
        of the reference's own code: tests/golden/refstubs.py.)"""
        raise NotImplementedError("CountRate.getCountRate has no host implementation here: the count rates are "
                                  "computed on the GPU (jx_xray); call fit.calcProfiles()")
