"""Test-side glue: turn a built ``fit`` object into the oracle's plain-array set-up."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import joxsz_oracle as orc  # noqa: E402


def oracle_setup_from_fit(fit):
    sz = fit.data.sz
    ann = fit.data.annuli
    ctr = ann.ctrate
    bands = []
    for b in fit.data.bands:
        t0, t1 = ctr.getTables(b.rmf, b.arf, b.emin_keV, b.emax_keV, fit.model.NH_1022pcm2)
        bands.append(dict(cts=np.asarray(b.cts, float), areascales=np.asarray(b.areascales, float),
                          exposures=np.asarray(b.exposures, float), backrates=np.asarray(b.backrates, float),
                          lnrate_Z0=np.asarray(t0, float), lnrate_Z1=np.asarray(t1, float)))
    names, kind, pa, pb, val = [], [], [], [], []
    for n, p in fit.pars.items():
        names.append(n)
        val.append(float(p.val))
        if hasattr(p, "prior_mu"):
            kind.append("gauss"); pa.append(float(p.prior_mu)); pb.append(float(p.prior_sigma))
        else:
            kind.append("box"); pa.append(float(p.minval)); pb.append(float(p.maxval))
    return orc.OracleSetup(
        phys_const=list(sz.phys_const), step=sz.step, kpc_as=sz.kpc_as,
        conv_T=np.asarray(sz.convert.x, float), conv_I=np.asarray(sz.convert.y, float),
        flux_data=np.asarray(sz.flux_data, float), beam_2d=np.asarray(sz.beam_2d, float),
        radius=np.asarray(sz.radius, float), sep=int(sz.sep), r_pp=np.asarray(sz.r_pp, float),
        d_mat=np.asarray(sz.d_mat, float), filtering=np.asarray(sz.filtering, float),
        calc_integ=bool(sz.calc_integ), integ_mu=sz.integ_mu, integ_sig=sz.integ_sig,
        midpt_kpc=np.asarray(ann.midpt_kpc, float), projvols_cm3=np.asarray(ann.projvols_cm3, float),
        geomarea_arcmin2=np.asarray(ann.geomarea_arcmin2, float), bands=bands,
        Tlogvals=np.asarray(ctr.Tlogvals, float), Tmin=float(ctr.Tmin), Tmax=float(ctr.Tmax),
        par_names=names, par_kind=kind, par_a=pa, par_b=pb, par_val=val, thawed=list(fit.thawed),
        dens_mode=fit.model.ne_cmpt.mode, exclude_unphy_mass=bool(fit.exclude_unphy_mass))


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / scale)


def rel_err_max(a, b):
    """max |a-b| / max |b| (for maps with zero crossings)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) / np.max(np.abs(b))
