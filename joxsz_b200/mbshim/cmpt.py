"""Profile components (``mbproj2.Cmpt/CmptFlat/CmptVikhDensity`` work-alikes)."""
import numpy as np

from .param import Param


class Cmpt:
    def __init__(self, name, annuli):
        self.name = name
        self.annuli = annuli

    def defPars(self):
        return {}

    def computeProf(self, pars):
        raise NotImplementedError

    def prior(self, pars):
        return 0.0


class CmptFlat(Cmpt):
    def __init__(self, name, annuli, defval=0.0, minval=-1e99, maxval=1e99):
        Cmpt.__init__(self, name, annuli)
        self.defval = defval
        self.minval = minval
        self.maxval = maxval

    def defPars(self):
        return {self.name: Param(self.defval, minval=self.minval, maxval=self.maxval)}

    def computeProf(self, pars):
        return np.full(self.annuli.nshells, float(pars[self.name].val))


class CmptVikhDensity(Cmpt):
    """Vikhlinin et al. (2006) density; JoXSZ overrides vikhFunction/defPars/prior
    (reference ``joxsz_main.py:135-139``), so only the plumbing lives here."""

    def __init__(self, name, annuli, mode="double"):
        Cmpt.__init__(self, name, annuli)
        self.mode = mode

    def vikhFunction(self, pars, radii_kpc):
        raise NotImplementedError("bind mydens_vikhFunction (joxsz_main.py:137)")

    def computeProf(self, pars):
        return self.vikhFunction(pars, self.annuli.midpt_kpc)

    def prior(self, pars):
        return 0.0
