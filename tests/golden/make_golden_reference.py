"""Generate golden vectors by running the reference's own code (see refstubs.py for what is stubbed).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_reference.py
Writes tests/golden/cl1226_inputs.npz   raw decoded contents of the reference's data files
       tests/golden/cl1226_golden.npz   reference outputs on seeded parameter draws

The set-up below calls the reference's functions in the order of joxsz_main.py:93-188 (the driver
itself cannot be imported: it starts the fit and the emcee run at import-free `main()` only, and
needs emcee/XSPEC).  Count-rate tables are the synthetic ones of joxsz_b200.synthetic (XSPEC absent).
"""
import os
import sys
from types import MethodType

import numpy as np
from scipy.interpolate import interp1d

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)

import refstubs  # noqa: E402


def build_reference_fit(ref, mb, data_dir):
    # joxsz_main.py:21-88 (configuration globals)
    mystep, m_e, sigma_T, R_b = 2., 0.5109989 * 1e3, 6.6524587158 * 1e-25, 5000.
    cosmology = mb.Cosmology(0.888)
    cosmology.H0, cosmology.WM, cosmology.WV = 67.32, 0.3158, 0.6842
    szd, xd = data_dir + "/SZ", data_dir + "/X"
    bandEs = [[700, 1000], [1000, 1300], [1300, 1600], [1600, 2000], [2000, 2700],
              [2700, 3400], [3400, 3800], [3800, 4300], [4300, 5000], [5000, 7000]]
    NH, Z_solar = 0.0183, 0.3
    rmf, arf = xd + "/source.rmf", xd + "/source.arf"
    infg, inbg = xd + "/fg_profnew_%04i_%04i.dat", xd + "/bg_profnew_%04i_%04i.dat"
    # joxsz_main.py:95-111
    phys_const = [m_e, sigma_T]
    kpc_as = cosmology.kpc_per_arcsec
    flux_data = ref.read_xy_err(szd + "/press_data_cl1226_flagsource_Xraycent.dat", ncol=3)
    maxr_data = flux_data[0][-1]
    beam_2d, fwhm = ref.mybeam(mystep, maxr_data, approx=False, filename=szd + "/Beam150GHz.fits",
                               normalize=True, fwhm_beam=None)
    mymaxr = (maxr_data + 3 * fwhm) // mystep * mystep
    radius = np.arange(0., mymaxr + mystep, mystep)
    radius = np.append(-radius[:0:-1], radius)
    sep = radius.size // 2
    r_pp = np.arange(mystep * kpc_as, R_b + mystep * kpc_as, mystep * kpc_as)
    d_mat = ref.centdistmat(radius * kpc_as)
    wn_as, tf = ref.read_tf(szd + "/TransferFunction150GHz_CLJ1227.fits", approx=False, loc=None, scale=None, c=None)
    filtering = ref.filt_image(wn_as, tf, d_mat.shape[0], mystep)
    t_keV, compt = np.loadtxt(szd + "/Compton_to_Jy_per_beam.dat", skiprows=1, unpack=True)
    convert = interp1d(t_keV, 1e3 * compt, 'linear', fill_value='extrapolate')
    sz_data = ref.SZ_data(phys_const, mystep, kpc_as, convert, flux_data, beam_2d, radius, sep, r_pp, d_mat,
                          filtering, False, .94 / 1e3, .36 / 1e3)
    # joxsz_main.py:116-125
    annuli = mb.Annuli(ref.getEdges(infg, bandEs), cosmology)
    bands = [ref.loadBand(infg, inbg, bandE, rmf, arf) for bandE in bandEs]
    data = mb.Data(bands, annuli)
    data.sz = sz_data
    # joxsz_main.py:128-175
    ref.add_param_unit()
    Z_cmpt = mb.CmptFlat('Z', annuli, defval=Z_solar, minval=0., maxval=1.)
    mb.CmptFlat.defPars = ref.Z_defPars
    ne_cmpt = mb.CmptVikhDensity('ne', annuli, mode='single')
    mb.CmptVikhDensity.vikhFunction = ref.mydens_vikhFunction
    mb.CmptVikhDensity.defPars = ref.mydens_defPars
    mb.CmptVikhDensity.prior = ref.mydens_prior
    press_cmpt = ref.CmptPressure('p', annuli)
    T_cmpt = ref.CmptUPPTemperature('T', annuli, press_cmpt, ne_cmpt)
    model = mb.ModelNullPot(annuli, ne_cmpt, T_cmpt, Z_cmpt, NH_1022pcm2=NH)
    pars = model.defPars()
    pars.update(press_cmpt.defPars())
    pars['backscale'] = mb.ParamGaussian(1., prior_mu=1., prior_sigma=0.1)
    pars['calibration'] = mb.ParamGaussian(1., prior_mu=1., prior_sigma=0.07)
    pars['log(r_c)'].maxval = annuli.edges_logkpc[-2]
    pars['log(r_s)'].maxval = annuli.edges_logkpc[-2]
    pars[r'\gamma'].val = 3.
    pars[r'\gamma'].frozen = True
    pars['log(r_c)'].val = 2.
    pars[r'\epsilon'].maxval = 10.
    pars[r'\alpha'].val = 0.
    pars[r'\alpha'].frozen = True
    pars['c'].frozen = True
    pars['log(T_X/T_{SZ})'].frozen = False
    # joxsz_main.py:178-188
    fit = mb.Fit(pars, model, data)
    fit.thawed = [name for name, par in fit.pars.items() if not par.frozen]
    fit.exclude_unphy_mass = True
    fit.savedir = "/tmp"
    fit.press = press_cmpt
    fit.mass_cmpt = ref.CmptMyMass('m', annuli, press_cmpt, ne_cmpt)
    mb.Fit.get_sz_like = MethodType(ref.get_sz_like, fit)
    mb.Fit.getLikelihood = MethodType(ref.getLikelihood, fit)
    mb.Fit.mylikeFromProfs = MethodType(ref.mylikeFromProfs, fit)
    return fit, bandEs


def main():
    """``--real-deps``: use the really installed PyAbel / mbproj2 / astropy instead of the restated stand-ins wherever
    they import (refstubs.real_deps_available), write ``cl1226_golden_realdeps.npz`` next to the committed goldens and
    print the largest difference of every recorded quantity against them -- the check that pins the third-party
    boundary, for a maintainer whose machine has those packages (the build container has none of them)."""
    from joxsz_b200 import cluster
    from joxsz_b200.synthetic import synthetic_countrate_tables, draw_parameters

    real_deps = "--real-deps" in sys.argv
    data_dir = os.path.join(refstubs.REFERENCE_DIR, "data")
    if real_deps:
        have = refstubs.real_deps_available()
        print("really installed:", have)
        if not (have["abel"] or have["mbproj2"]):
            raise SystemExit("--real-deps: neither PyAbel nor mbproj2 is importable here; nothing to pin")
    else:
        # 1. raw inputs fixture (this repo's readers; checked against the reference's readers below)
        inp = cluster.load_cl1226_files(data_dir)
        cluster.save_inputs_npz(inp, os.path.join(HERE, "cl1226_inputs.npz"))

    # 2. the reference's own code
    ref, mb = refstubs.import_reference(real_deps)
    mb.fit.debugfit = False
    fit, bandEs = build_reference_fit(ref, mb, data_dir)
    ctr = fit.data.annuli.ctrate
    tables = synthetic_countrate_tables([(b.emin_keV, b.emax_keV) for b in fit.data.bands], ctr.Tlogvals)
    for band, (t0, t1) in zip(fit.data.bands, tables):
        # key layout documented by the reference's own addCountCache (joxsz_funcs.py:652-681)
        key = (band.emin_keV, band.emax_keV, ctr.cosmo.z, fit.model.NH_1022pcm2, band.rmf, band.arf)
        ctr.ctcache[key] = (np.asarray(t0, dtype=np.float64), np.asarray(t1, dtype=np.float64))

    thawed = list(fit.thawed)
    default_theta = np.array(fit.thawedParVals(), dtype=np.float64)
    draws = draw_parameters(thawed, n=47, seed=20260101, frac_bad=0.15)
    thetas = np.vstack([default_theta[None, :], draws])
    W = thetas.shape[0]
    sz = fit.data.sz
    nb, na = len(fit.data.bands), fit.data.annuli.nshells
    out = dict(ll=np.empty(W), pp=np.empty((W, sz.r_pp.size)), bright=np.empty((W, sz.sep + 1)),
               chisq=np.empty(W), szll=np.empty(W), xprofs=np.empty((W, nb, na)), xlike=np.empty(W),
               mass=np.empty((W, sz.r_pp.size)), tsz=np.empty((W, sz.sep)))
    with np.errstate(all="ignore"):
        for w in range(W):
            out["ll"][w] = fit.getLikelihood(thetas[w])          # also leaves fit.pars at thetas[w]
            out["pp"][w] = fit.get_sz_like(output="pp")
            out["bright"][w] = fit.get_sz_like(output="bright")
            out["chisq"][w] = fit.get_sz_like(output="chisq")
            out["szll"][w] = fit.get_sz_like(output="ll")
            profs = fit.calcProfiles()
            out["xprofs"][w] = np.array(profs)
            out["xlike"][w] = fit.mylikeFromProfs(profs) if np.array(profs).min() > 0 else -np.inf
            out["mass"][w] = fit.mass_cmpt.mass_fun(fit.pars, sz.r_pp)
            out["tsz"][w] = fit.model.T_cmpt.temp_fun(fit.pars, sz.r_pp[:sz.sep], getT_SZ=True)
    # second pass with the integrated-Compton-parameter penalty switched on (joxsz_funcs.py:480-487;
    # `simps` is scipy's `simpson` of the installed version, see refstubs.py)
    sz.calc_integ = True
    out.update(ll_integ=np.empty(W), szll_integ=np.empty(W), cint=np.empty(W))
    with np.errstate(all="ignore"):
        for w in range(W):
            out["ll_integ"][w] = fit.getLikelihood(thetas[w])
            out["szll_integ"][w] = fit.get_sz_like(output="ll")
            out["cint"][w] = fit.get_sz_like(output="integ")
    sz.calc_integ = False
    print("finite ll:", int(np.isfinite(out["ll"]).sum()), "of", W, " ll[0] =", out["ll"][0])
    outname = "cl1226_golden_realdeps.npz" if real_deps else "cl1226_golden.npz"
    if real_deps:
        old = np.load(os.path.join(HERE, "cl1226_golden.npz"))
        print("largest differences against the committed goldens (third-party stand-ins):")
        for k, v in out.items():
            if k in old.files:
                a, b = np.asarray(v, float), np.asarray(old[k], float)
                fin = np.isfinite(a) & np.isfinite(b)
                same_mask = bool(np.array_equal(np.isfinite(a), np.isfinite(b)))
                scale = np.max(np.abs(b[fin])) if fin.any() else 1.0
                print(f"  {k:12s} max |d| = {np.max(np.abs(a[fin] - b[fin])) if fin.any() else 0.0:.3e}"
                      f"  (/ max |golden| = {np.max(np.abs(a[fin] - b[fin])) / scale if fin.any() else 0.0:.3e})"
                      f"  finite mask equal: {same_mask}")
    np.savez_compressed(
        os.path.join(HERE, outname), deps_mode=np.array(sorted(f"{k}={v}" for k, v in refstubs.MODE.items())),
        thawed=np.array(thawed), thetas=thetas,
        r_pp=sz.r_pp, radius=sz.radius, sep=np.array(sz.sep), kpc_as=np.array(sz.kpc_as),
        beam_2d=sz.beam_2d, filtering=sz.filtering, d_mat_row=sz.d_mat[sz.sep],
        midpt_kpc=fit.data.annuli.midpt_kpc, projvols_cm3=fit.data.annuli.projvols_cm3,
        par_names=np.array(list(fit.pars.keys())), integ_mu=np.array(sz.integ_mu), integ_sig=np.array(sz.integ_sig),
        scipy_version=np.array(__import__('scipy').__version__), **out)


if __name__ == "__main__":
    main()
