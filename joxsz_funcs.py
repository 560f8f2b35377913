"""Drop-in for the reference's ``joxsz_funcs`` module: every name ``joxsz_main.py:10-12`` imports,
served by the B200 implementation in :mod:`joxsz_b200`."""
from joxsz_b200 import (SZ_data, read_xy_err, mybeam, centdistmat, read_tf, filt_image, getEdges, loadBand,  # noqa: F401
                        add_param_unit, Z_defPars, CmptPressure, CmptUPPTemperature, CmptMyMass, mydens_defPars,
                        mydens_vikhFunction, mydens_prior, get_sz_like, mylikeFromProfs, getLikelihood,
                        add_backend_attrs, addCountCache, read_beam, dist, calcProfiles)


def mcmc_run(*args, **kwargs):
    from joxsz_b200.sampler import mcmc_run as _run
    return _run(*args, **kwargs)
