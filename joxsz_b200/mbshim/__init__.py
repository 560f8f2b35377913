"""Minimal ``mbproj2`` work-alike: only what JoXSZ's driver and likelihood touch.

mbproj2 is an unpinned, un-vendored dependency of the reference (``requirements.txt``) and is
not installable in this environment; XSPEC (its table builder) is absent too.  The classes here
restate its public behaviour from memory so that the reference's
``joxsz_main.py`` flow -- build annuli/bands/model/pars/fit, bind the JoXSZ methods -- can be
reproduced.  When the real mbproj2 is importable it is used instead (see ``joxsz_b200.mb``).
All of this is host-side set-up; none of it is on the batched likelihood path.
"""
from . import physconstants, utils, xspechelper, countrate, fit  # noqa: F401
from .cosmo import Cosmology  # noqa: F401
from .param import ParamBase, Param, ParamGaussian  # noqa: F401
from .cmpt import Cmpt, CmptFlat, CmptVikhDensity  # noqa: F401
from .data import Annuli, Band, Data  # noqa: F401
from .model import ModelNullPot  # noqa: F401
from .fit import Fit  # noqa: F401
