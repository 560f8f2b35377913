"""Parameter objects (``mbproj2.ParamBase/Param/ParamGaussian`` work-alikes).

``prior()`` semantics: a box parameter contributes ``-inf``
outside ``[minval, maxval]`` and 0 inside; a Gaussian parameter contributes the normal
log-pdf.  JoXSZ replaces both ``__init__`` methods to add ``unit``
(reference ``joxsz_funcs.py:213-239``) and leaves ``prior()`` untouched.
"""
import math

import numpy as np


class ParamBase:
    def __init__(self, val, frozen=False):
        self.val = val
        self.frozen = frozen

    def prior(self):
        return 0.0

    def copy(self):
        import copy
        return copy.copy(self)


class Param(ParamBase):
    def __init__(self, val, minval=-1e99, maxval=1e99, frozen=False):
        ParamBase.__init__(self, val, frozen=frozen)
        self.minval = minval
        self.maxval = maxval

    def prior(self):
        if self.val < self.minval or self.val > self.maxval:
            return -np.inf
        return 0.0

    def __repr__(self):
        return "<Param: val=%.3g, minval=%.3g, maxval=%.3g, frozen=%s>" % (
            self.val, self.minval, self.maxval, self.frozen)


class ParamGaussian(ParamBase):
    def __init__(self, val, prior_mu, prior_sigma, frozen=False):
        ParamBase.__init__(self, val, frozen=frozen)
        self.prior_mu = prior_mu
        self.prior_sigma = prior_sigma

    def prior(self):
        if self.prior_sigma <= 0:
            return 0.0
        return (-0.5 * math.log(2.0 * math.pi) - math.log(self.prior_sigma)
                - 0.5 * ((self.val - self.prior_mu) / self.prior_sigma) ** 2)

    def __repr__(self):
        return "<ParamGaussian: val=%.3g, prior_mu=%.3g, prior_sigma=%.3g, frozen=%s>" % (
            self.val, self.prior_mu, self.prior_sigma, self.frozen)
