"""joxsz_b200 -- the JoXSZ joint SZ + X-ray log-likelihood, batched over every emcee walker, on B200.

Public surface (mirrors what ``joxsz_main.py`` imports from ``joxsz_funcs``; see the top-level
``joxsz_funcs.py`` shim): readers, ``SZ_data``, the profile components, ``get_sz_like`` /
``mylikeFromProfs`` / ``getLikelihood`` to bind onto ``mbproj2.Fit``, ``mcmc_run``; plus
``BatchedLikelihood`` (theta[W, ndim] -> ll[W]) and the device ensemble sampler.

Compute lives in ``libjoxsz_b200.so`` (hand-written sm_100a CUDA, C ABI in ``include/joxsz_b200.h``).
Importing this package does not need a GPU; evaluating anything does.
"""
from .readers import (SZ_data, read_xy_err, read_beam, mybeam, centdistmat, read_tf, dist, filt_image,  # noqa: F401
                      getEdges, loadBand)
from .components import (add_param_unit, Z_defPars, CmptPressure, CmptUPPTemperature, CmptMyMass,  # noqa: F401
                         mydens_defPars, mydens_vikhFunction, mydens_prior)
from .funcs import (get_sz_like, mylikeFromProfs, getLikelihood, calcProfiles, add_backend_attrs,  # noqa: F401
                    addCountCache, engine_for)
from .mb import mb, USING_SHIM  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch-dependent pieces are imported lazily so that host-only tooling stays light
    if name == "BatchedLikelihood":
        from .batched import BatchedLikelihood
        return BatchedLikelihood
    if name in ("EnsembleSampler", "mcmc_run"):
        from . import sampler
        return getattr(sampler, name)
    raise AttributeError(name)
