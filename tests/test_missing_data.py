"""Edge cases of the data the reference tolerates: NaN counts in X-ray bins (masked out of the Cash sum,
reference joxsz_funcs.py:504), NaN SZ points (dropped by the nansum at :478), zero-count bins, and the
mass veto switched off (joxsz_main.py:88 `exclude_unphy_mass = False`)."""
import os

import numpy as np
import pytest

from helpers import oracle_setup_from_fit, orc

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _build(**changes):
    from joxsz_b200 import cluster
    from joxsz_b200.mb import mb
    mb.fit.debugfit = False
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    xfg = np.array(inp.xfg, dtype=np.float64)
    flux = np.array(inp.flux_data, dtype=np.float64)
    if changes.get("nan_counts"):
        xfg[0, 3, 2] = np.nan          # band 0, annulus 3
        xfg[4, :2, 2] = np.nan         # band 4, two inner annuli
        xfg[9, 14, 2] = np.nan         # last band, outermost annulus
    if changes.get("zero_counts"):
        xfg[2, :, 2] = 0.0             # a band with no photons at all
    if changes.get("nan_flux"):
        flux[1, 5] = np.nan            # one SZ point missing
        flux[2, 11] = np.nan           # one SZ error missing
    inp.xfg, inp.flux_data = xfg, flux
    if "exclude_unphy_mass" in changes:
        inp.exclude_unphy_mass = changes["exclude_unphy_mass"]
    fit, _ = cluster.build_fit(inp, savedir=None)
    return fit


@pytest.fixture(scope="module")
def fit_gappy():
    return _build(nan_counts=True, zero_counts=True, nan_flux=True)


@pytest.fixture(scope="module")
def fit_noveto():
    return _build(exclude_unphy_mass=False)


def test_oracle_masks_missing_bins(fit_gappy, cl1226_oracle, golden):
    s = oracle_setup_from_fit(fit_gappy)
    th = golden["thetas"][:12]
    a = orc.get_likelihood_many(th, s)
    b = orc.get_likelihood_many(th, cl1226_oracle)
    ok = np.isfinite(a)
    assert np.array_equal(ok, np.isfinite(b)) and ok.sum() >= 5
    assert np.all(np.abs(a[ok] - b[ok]) > 1e-3)          # the masked bins really changed the likelihood


@pytest.mark.gpu
def test_gpu_with_missing_bins(fit_gappy, golden):
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.synthetic import draw_parameters
    s = oracle_setup_from_fit(fit_gappy)
    eng = BatchedLikelihood(fit_gappy, max_walkers=256)
    th = np.vstack([golden["thetas"], draw_parameters(fit_gappy.thawed, n=80, seed=12, spread=0.03, frac_bad=0.1)])
    ll = eng(th)
    ref = orc.get_likelihood_many(th, s)
    assert not np.isnan(ll).any()
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert ok.sum() > 40
    assert np.max(np.abs(ll[ok] - ref[ok])) < 1e-6
    # the taps agree on the masked Cash sum and on the nansum chi^2
    x = eng.xray(th[:6])
    z = eng.sz_profile(th[:6])
    for k in range(6):
        p = s.full_params(th[k])
        with np.errstate(all="ignore"):
            profs = orc.xray_profiles(p, s)
            st = orc.sz_stages(p, s)
        if np.array(profs).min() > 0:
            assert abs(x["cash"][k] - orc.xray_like_from_profs(profs, s)) < 1e-6
        if np.isfinite(st["chisq"]):
            assert abs(z["chisq"][k] - st["chisq"]) < 1e-6
    eng.close()


@pytest.mark.gpu
def test_gpu_without_mass_veto(fit_noveto, golden, cl1226_oracle):
    from joxsz_b200.batched import BatchedLikelihood
    s = oracle_setup_from_fit(fit_noveto)
    assert not s.exclude_unphy_mass
    eng = BatchedLikelihood(fit_noveto, max_walkers=256)
    th = golden["thetas"]
    ll = eng(th)
    ref = orc.get_likelihood_many(th, s)
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert np.max(np.abs(ll[ok] - ref[ok])) < 1e-6
    # some walkers the veto rejects are finite now
    assert ok.sum() > np.isfinite(golden["ll"]).sum()
    eng.close()
