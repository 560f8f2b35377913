"""GPU parity: every stage of libjoxsz_b200.so, called through the C ABI, against the CPU oracle and
the golden vectors recorded from the reference's own code.

Tolerances are the ones BASELINE.json states: per-walker log-likelihood within 1e-6 ABSOLUTE, model
maps and profiles within 1e-5 RELATIVE; the -inf mask must agree exactly.  (Measured agreement is far
tighter; the asserts below also enforce a 1e-9 relative regression guard on the smooth quantities.)
"""
import numpy as np
import pytest
import torch

from helpers import orc, rel_err, rel_err_max

pytestmark = pytest.mark.gpu

LL_ATOL = 1e-6
PROF_RTOL = 1e-5


@pytest.fixture(scope="module")
def engine(cl1226_fit):
    from joxsz_b200.batched import BatchedLikelihood
    eng = BatchedLikelihood(cl1226_fit, max_walkers=4096)
    yield eng
    eng.close()


def test_profiles_match_reference_golden(engine, golden):
    th = golden["thetas"]
    out = engine.profiles(th)
    assert rel_err(out["pp"], golden["pp"]) < PROF_RTOL
    assert rel_err(out["pp"], golden["pp"]) < 1e-12
    assert rel_err(out["tsz"], golden["tsz"]) < 1e-12
    # status bits reproduce the reference's -inf mask
    assert np.array_equal(out["flags"] != 0, ~np.isfinite(golden["ll"]))


def test_component_methods_match_oracle(cl1226_fit, cl1226_oracle, golden):
    s = cl1226_oracle
    fit = cl1226_fit
    r = np.geomspace(5.0, 4000.0, 77)
    saved = fit.thawedParVals()
    try:
        for w in (0, 3, 11):
            fit.updateThawed(golden["thetas"][w])
            p = s.full_params(golden["thetas"][w])
            assert rel_err(fit.press.press_fun(fit.pars, r), orc.press_fun(p, r)) < 1e-13
            assert rel_err(fit.press.press_derivative(fit.pars, r), orc.press_derivative(p, r)) < 1e-13
            assert rel_err(fit.model.ne_cmpt.vikhFunction(fit.pars, r), orc.vikh_density(p, r)) < 1e-13
            assert rel_err(fit.model.T_cmpt.temp_fun(fit.pars, r), orc.temp_fun(p, r)) < 1e-13
            assert rel_err(fit.model.T_cmpt.temp_fun(fit.pars, r, getT_SZ=True), orc.temp_fun(p, r, getT_SZ=True)) < 1e-13
            assert rel_err(fit.mass_cmpt.mass_fun(fit.pars, r), orc.mass_fun(p, r)) < 1e-12
    finally:
        fit.updateThawed(saved)


def test_projection_matches_oracle(engine, cl1226_oracle, golden):
    s = cl1226_oracle
    idx = [0, 1, 4, 9, 20]
    out = engine.sz_project(golden["thetas"][idx])
    for k, w in enumerate(idx):
        with np.errstate(all="ignore"):
            st = orc.sz_stages(s.full_params(golden["thetas"][w]), s)
        assert rel_err_max(out["y"][k], st["y"]) < 1e-12
        # the spline pieces reproduce scipy's interp1d on the map
        mo = engine.packed.map_ops
        c = out["coef"][k]
        Z = c[0][mo.seg] + mo.dx * (c[1][mo.seg] + mo.dx * (c[2][mo.seg] + mo.dx * c[3][mo.seg]))
        cc = s.d_mat.shape[0] // 2
        assert rel_err_max(Z, st["y_2d"][cc:, cc:]) < 1e-11


def test_maps_match_oracle(engine, cl1226_oracle, golden):
    s = cl1226_oracle
    idx = [0, 2, 7]
    maps = engine.sz_maps(golden["thetas"][idx])
    for k, w in enumerate(idx):
        with np.errstate(all="ignore"):
            st = orc.sz_stages(s.full_params(golden["thetas"][w]), s)
        for name in ("y_2d", "conv_2d", "map_out"):
            err = rel_err_max(maps[name][k], st[name])
            assert err < PROF_RTOL, (name, err)
            assert err < 1e-10, (name, err)


def test_sz_profile_matches_reference_golden(engine, golden):
    th = golden["thetas"]
    out = engine.sz_profile(th)
    ok = np.isfinite(golden["chisq"]) & ~np.isnan(golden["bright"]).any(axis=1)
    scale = np.max(np.abs(golden["bright"][ok]), axis=1, keepdims=True)
    err = np.max(np.abs(out["bright"][ok] - golden["bright"][ok]) / scale)
    assert err < PROF_RTOL and err < 1e-10, err
    dchi = np.abs(out["chisq"][ok] - golden["chisq"][ok])
    assert np.max(dchi) < LL_ATOL, np.max(dchi)
    # the production row equals the full-map tap's central row
    maps = engine.sz_maps(th[:2], want=("map_out",))
    c = engine.packed.N // 2
    assert rel_err_max(out["row"][:2], maps["map_out"][:, c, c:]) < 1e-11


def test_xray_matches_reference_golden(engine, golden):
    th = golden["thetas"]
    out = engine.xray(th)
    ok = np.isfinite(golden["xprofs"]).all(axis=(1, 2))
    assert rel_err(out["pred"][ok], golden["xprofs"][ok]) < 1e-12
    fin = np.isfinite(golden["xlike"])
    assert np.array_equal(np.isfinite(out["cash"]), fin)
    assert np.max(np.abs(out["cash"][fin] - golden["xlike"][fin])) < LL_ATOL
    # mylikeFromProfs on supplied profiles
    got = engine.cash_from_profiles(golden["xprofs"][fin])
    assert np.max(np.abs(got - golden["xlike"][fin])) < LL_ATOL


def test_loglike_matches_reference_golden(engine, golden):
    th = golden["thetas"]
    ll = engine(th)
    ref = golden["ll"]
    assert not np.isnan(ll).any()
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert np.max(np.abs(ll[ok] - ref[ok])) < LL_ATOL, np.max(np.abs(ll[ok] - ref[ok]))
    # single-vector call returns a float, CUDA tensor in gives CUDA tensor out
    assert abs(engine(th[0]) - ref[0]) < LL_ATOL
    t = torch.from_numpy(th).cuda()
    out = engine(t)
    assert out.is_cuda and torch.equal(out.cpu(), torch.from_numpy(ll))


def test_loglike_1024_walkers_vs_oracle(engine, cl1226_fit, cl1226_oracle):
    """BASELINE config 2: 1024 synthetic draws, per-walker check against the oracle."""
    from joxsz_b200.synthetic import draw_parameters
    s = cl1226_oracle
    th = draw_parameters(s.thawed, n=1024, seed=20260102, frac_bad=0.05)
    ll = engine(th)
    ref = orc.BatchedOracle(s).loglike(th)
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert ok.sum() > 500
    assert np.max(np.abs(ll[ok] - ref[ok])) < LL_ATOL, np.max(np.abs(ll[ok] - ref[ok]))
    # literal per-walker path on a subset
    sub = np.arange(0, 1024, 64)
    lit = orc.get_likelihood_many(th[sub], s)
    fin = np.isfinite(lit)
    assert np.array_equal(fin, np.isfinite(ll[sub]))
    assert np.max(np.abs(ll[sub][fin] - lit[fin])) < LL_ATOL


def test_properties_at_full_batch(engine):
    """Size-independent properties at a batch far larger than the oracle could check walker by walker:
    permutation equivariance, independence of the batch composition, determinism."""
    from joxsz_b200.synthetic import draw_parameters
    W = 4096
    th = draw_parameters(engine.packed.thawed, n=W, seed=99, frac_bad=0.1)
    ll = engine(th)
    assert not np.isnan(ll).any()
    perm = np.random.default_rng(0).permutation(W)
    assert np.array_equal(engine(th[perm]), ll[perm])
    assert np.array_equal(engine(th[:1000]), ll[:1000])
    # ragged batches: a last half-warp without a walker (X-ray kernel), a last GEMM tile with 1 / 127 / 129 rows
    for n in (1, 127, 129, 999):
        assert np.array_equal(engine(th[7:7 + n]), ll[7:7 + n]), n
    assert np.array_equal(engine(th), ll)
    # linearity of the SZ chain in P_0 (calibration fixed): bright scales with P_0 at fixed shape... T_SZ
    # also scales, so test the filtered row, which is linear in the pressure profile
    good = th[np.isfinite(ll)][:8].copy()
    j = engine.packed.thawed.index("P_0")
    r1 = engine.sz_profile(good)["row"]
    good2 = good.copy(); good2[:, j] *= 0.5
    r2 = engine.sz_profile(good2)["row"]
    assert rel_err_max(r2, 0.5 * r1) < 1e-12


def test_edge_cases(engine, golden):
    th = golden["thetas"]
    assert engine(th[:0].reshape(0, th.shape[1])).shape == (0,)
    with pytest.raises(ValueError):
        engine(np.zeros((3, th.shape[1] + 1)))
    with pytest.raises(ValueError):
        engine(np.zeros((engine.max_walkers + 1, th.shape[1])))
    # NaN parameters never produce NaN log-likelihoods (emcee would raise on NaN)
    bad = th[:4].copy()
    bad[0, 0] = np.nan; bad[1, 7] = np.nan; bad[2, 11] = np.nan; bad[3, 12] = np.inf
    out = engine(bad)
    assert not np.isnan(out).any() and np.all(out == -np.inf)


def test_getLikelihood_dropin(cl1226_fit, golden):
    """The reference-facing method bound on Fit (joxsz_main.py:187): scalar and batched calls."""
    fit = cl1226_fit
    saved = fit.thawedParVals()
    try:
        v = fit.getLikelihood(golden["thetas"][0])
        assert isinstance(v, float) and abs(v - golden["ll"][0]) < LL_ATOL
        assert fit.thawedParVals() == list(golden["thetas"][0])
        assert abs(fit.getLikelihood() - golden["ll"][0]) < LL_ATOL
        out = fit.getLikelihood(golden["thetas"][:8])
        ok = np.isfinite(golden["ll"][:8])
        assert np.max(np.abs(out[ok] - golden["ll"][:8][ok])) < LL_ATOL
        assert abs(fit.get_sz_like() - golden["szll"][0]) < LL_ATOL
        assert rel_err(fit.get_sz_like(output="pp"), golden["pp"][0]) < 1e-12
        assert rel_err_max(fit.get_sz_like(output="bright"), golden["bright"][0]) < 1e-10
        profs = fit.calcProfiles()
        assert rel_err(np.array(profs), golden["xprofs"][0]) < 1e-12
        assert abs(fit.mylikeFromProfs(profs) - golden["xlike"][0]) < LL_ATOL
    finally:
        fit.updateThawed(saved)


def test_fft_form_of_the_beam_convolution(cl1226_fit, golden, engine, monkeypatch):
    """The map kernel convolves along y directly when the beam is small (the shipped case) and through two column
    FFTs otherwise; JX_K3_BFFT=1 forces the FFT form so that both stay covered.  Same maps, same likelihood."""
    from joxsz_b200.batched import BatchedLikelihood
    monkeypatch.setenv("JX_K3_BFFT", "1")
    eng = BatchedLikelihood(cl1226_fit, max_walkers=256)
    try:
        th = golden["thetas"]
        ll_fft = eng(th)
        maps_fft = eng.sz_maps(th[:2], want=("conv_2d",))["conv_2d"]
    finally:
        eng.close()
    monkeypatch.delenv("JX_K3_BFFT")
    ll = engine(th)
    ok = np.isfinite(golden["ll"])
    assert np.array_equal(np.isfinite(ll_fft), ok)
    assert np.max(np.abs(ll_fft[ok] - golden["ll"][ok])) < LL_ATOL
    assert np.max(np.abs(ll_fft[ok] - ll[ok])) < 1e-9
    maps = engine.sz_maps(th[:2], want=("conv_2d",))["conv_2d"]
    assert rel_err_max(maps_fft, maps) < 1e-12


def test_pinned_host_input(engine, golden):
    """A page-locked float64 torch tensor is copied to the device directly (no staging copy) and gives the same
    values as the numpy path; the result comes back as a host tensor."""
    th = golden["thetas"]
    ref = engine(th)
    pinned = torch.from_numpy(np.ascontiguousarray(th)).pin_memory()
    out = engine(pinned)
    assert isinstance(out, torch.Tensor) and not out.is_cuda
    assert np.array_equal(out.numpy(), ref)


@pytest.mark.gpu
def test_collapsed_mode_matches_staged_and_golden(cl1226_fit, golden):
    """Optional `ll`-only mode: the linear SZ chain folded into one operator built from the staged kernels themselves
    (jx_loglike_collapsed).  Same -inf mask, log-likelihoods within 1e-8 of the staged path and 1e-6 of the
    reference-recorded goldens."""
    from joxsz_b200.batched import BatchedLikelihood
    thetas = golden["thetas"]
    staged = BatchedLikelihood(cl1226_fit, max_walkers=256)
    coll = BatchedLikelihood(cl1226_fit, max_walkers=100, mode="collapsed")      # < nr walkers: operator built in chunks
    a, b = staged(thetas), coll(thetas)
    fin = np.isfinite(golden["ll"])
    assert np.array_equal(np.isfinite(a), fin) and np.array_equal(np.isfinite(b), fin)
    assert np.max(np.abs(a[fin] - b[fin])) < 1e-8
    assert np.max(np.abs(b[fin] - golden["ll"][fin])) < 1e-6
    assert not np.isnan(b).any()
    staged.close(); coll.close()


@pytest.mark.gpu
def test_loglike_bits_do_not_depend_on_the_batch(cl1226_fit):
    """A walker's log-likelihood is bit-identical whatever batch it is evaluated in -- sizes around the persistent grid
    of the map kernel (148 CTAs, two walkers in flight each), with skipped (flagged) walkers in between -- and equal to the
    K3 form of the map stage (JX_K3_WS=0 engine).  The N-GPU chain is the 1-GPU chain only because of this."""
    import os
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.synthetic import draw_parameters
    theta = draw_parameters(cl1226_fit.thawed, n=600, seed=31, spread=0.03, frac_bad=0.2)
    eng = BatchedLikelihood(cl1226_fit, max_walkers=600)
    full = eng(theta)
    assert np.isinf(full).sum() > 50 and np.isfinite(full).sum() > 300
    for n in (1, 2, 3, 147, 148, 149, 295, 296, 297, 445, 599):
        part = eng(theta[:n])
        assert np.array_equal(part.view(np.int64), full[:n].view(np.int64)), n
        tail = eng(theta[600 - n:])
        assert np.array_equal(tail.view(np.int64), full[600 - n:].view(np.int64)), n
    old = os.environ.get("JX_K3_WS")
    os.environ["JX_K3_WS"] = "0"
    try:
        eng0 = BatchedLikelihood(cl1226_fit, max_walkers=600)
    finally:
        if old is None:
            del os.environ["JX_K3_WS"]
        else:
            os.environ["JX_K3_WS"] = old
    assert np.array_equal(eng0(theta).view(np.int64), full.view(np.int64))
    eng.close(); eng0.close()
