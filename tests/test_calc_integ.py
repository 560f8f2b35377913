"""Row S8 of the scope table: the optional integrated-Compton-parameter penalty
(reference joxsz_funcs.py:480-487, `calc_integ`, off by default at joxsz_main.py:65).

Golden vectors `cint`, `ll_integ`, `szll_integ` were recorded from the reference's own get_sz_like /
getLikelihood with calc_integ = True (tests/golden/make_golden_reference.py); its `simps` is the installed
scipy's Simpson rule, so these fixtures pin the behaviour for the scipy version stored next to them."""
import numpy as np
import pytest

from helpers import oracle_setup_from_fit, orc


def test_oracle_matches_reference_with_calc_integ(golden, cl1226_fit_integ):
    s = oracle_setup_from_fit(cl1226_fit_integ)
    assert s.calc_integ
    th = golden["thetas"][:24]
    ll = orc.get_likelihood_many(th, s)
    ref = golden["ll_integ"][:24]
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert np.max(np.abs(ll[ok] - ref[ok])) < 1e-9
    for w in (0, 1, 5):
        with np.errstate(all="ignore"):
            st = orc.sz_stages(s.full_params(th[w]), s)
        assert abs(st["integ"] - golden["cint"][w]) <= 1e-13 * abs(golden["cint"][w])
        assert abs(st["ll"] - golden["szll_integ"][w]) < 1e-9


def test_integ_operator_reproduces_reference_cint(golden, cl1226_fit_integ):
    """The packed linear functional w_integ (Simpson * 2 pi * y scaling * Abel) applied to the reference's
    own pressure profiles gives the reference's cint."""
    from joxsz_b200.packer import PackedSetup
    pk = PackedSetup(cl1226_fit_integ, max_walkers=8)
    assert pk.calc_integ and pk.w_integ.shape == (pk.nr,)
    pp, cint = golden["pp"], golden["cint"]
    ok = np.isfinite(cint) & np.isfinite(pp).all(axis=1)
    got = pp[ok] @ pk.w_integ
    assert np.max(np.abs(got - cint[ok]) / np.abs(cint[ok])) < 1e-12


@pytest.mark.gpu
def test_gpu_loglike_with_calc_integ(golden, cl1226_fit_integ):
    from joxsz_b200.batched import BatchedLikelihood
    eng = BatchedLikelihood(cl1226_fit_integ, max_walkers=256)
    th = golden["thetas"]
    ll = eng(th)
    ref = golden["ll_integ"]
    assert not np.isnan(ll).any()
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert np.max(np.abs(ll[ok] - ref[ok])) < 1e-6
    # the penalty is really in there: it differs from the calc_integ = False likelihood
    assert np.max(np.abs(golden["ll"][ok] - ref[ok])) > 1e-3
    out = eng.sz_profile(th)
    fin = np.isfinite(golden["cint"])
    assert np.max(np.abs(out["cint"][fin] - golden["cint"][fin]) / np.abs(golden["cint"][fin])) < 1e-11
    eng.close()


@pytest.mark.gpu
def test_get_sz_like_integ_output(golden, cl1226_fit_integ, cl1226_fit):
    fit = cl1226_fit_integ
    saved = fit.thawedParVals()
    try:
        fit.updateThawed(golden["thetas"][0])
        assert abs(fit.get_sz_like(output="integ") - golden["cint"][0]) < 1e-11 * abs(golden["cint"][0])
        assert abs(fit.get_sz_like() - golden["szll_integ"][0]) < 1e-6
        assert abs(fit.get_sz_like(output="chisq") - golden["chisq"][0]) < 1e-6
        assert abs(fit.getLikelihood(golden["thetas"][0]) - golden["ll_integ"][0]) < 1e-6
    finally:
        fit.updateThawed(saved)
    with pytest.raises(RuntimeError):            # calc_integ off: 'integ' is not a valid output (reference :492)
        cl1226_fit.get_sz_like(output="integ")
    # two fits in one process stay independent
    assert abs(cl1226_fit.getLikelihood(golden["thetas"][0]) - golden["ll"][0]) < 1e-6
