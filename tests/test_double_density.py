"""Density mode 'double' (second beta-model term, reference joxsz_funcs.py:390-394): no shipped
configuration uses it, so parity is checked against the oracle on seeded draws."""
import os

import numpy as np
import pytest

from helpers import oracle_setup_from_fit, orc

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def fit_double():
    from joxsz_b200 import cluster
    from joxsz_b200.mb import mb
    mb.fit.debugfit = False
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    inp.dens_mode = "double"
    fit, _ = cluster.build_fit(inp, savedir=None)
    return fit


def _draws(fit, n, seed):
    from joxsz_b200.synthetic import FIDUCIAL, draw_parameters
    fid = dict(FIDUCIAL)
    fid.update({"log(n_{02})": -2.0, r"\beta_2": 0.9, "log(r_{c2})": 1.6})       # keeps the HSE mass monotone
    return draw_parameters(fit.thawed, fiducial=fid, n=n, seed=seed, spread=0.03, frac_bad=0.1)


def test_double_mode_has_three_more_parameters(fit_double, cl1226_fit):
    assert set(fit_double.thawed) - set(cl1226_fit.thawed) == {"log(n_{02})", r"\beta_2", "log(r_{c2})"}
    s = oracle_setup_from_fit(fit_double)
    assert s.dens_mode == "double"
    th = _draws(fit_double, 8, 3)
    ll = orc.get_likelihood_many(th, s)
    assert np.isfinite(ll).sum() >= 3


@pytest.mark.gpu
def test_double_mode_gpu_matches_oracle(fit_double):
    from joxsz_b200.batched import BatchedLikelihood
    s = oracle_setup_from_fit(fit_double)
    eng = BatchedLikelihood(fit_double, max_walkers=256)
    th = _draws(fit_double, 96, 5)
    ll = eng(th)
    ref = orc.get_likelihood_many(th, s)
    assert not np.isnan(ll).any()
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert ok.sum() > 30
    assert np.max(np.abs(ll[ok] - ref[ok])) < 1e-6
    prof = eng.profiles(th[:8])
    for k in range(8):
        p = s.full_params(th[k])
        ne = orc.vikh_density(p, s.midpt_kpc, "double")
        assert np.max(np.abs(prof["ne_ann"][k] - ne) / ne) < 1e-12
    eng.close()
