"""Pin the oracle against outputs of the reference's OWN code (tests/golden/make_golden_reference.py).

What this pins and what it does not is spelled out in oracle/joxsz_oracle.py: the reference's
joxsz_funcs.py logic is pinned; PyAbel / mbproj2 internals are stubs shared with the oracle.
"""
import numpy as np
import pytest

from helpers import orc, rel_err


def test_setup_matches_reference_geometry(golden, cl1226_oracle):
    s = cl1226_oracle
    assert list(golden["thawed"]) == s.thawed
    assert list(golden["par_names"]) == s.par_names
    assert int(golden["sep"]) == s.sep
    np.testing.assert_array_equal(golden["r_pp"], s.r_pp)
    np.testing.assert_array_equal(golden["radius"], s.radius)
    np.testing.assert_array_equal(golden["beam_2d"], s.beam_2d)
    np.testing.assert_array_equal(golden["filtering"], s.filtering)
    np.testing.assert_array_equal(golden["d_mat_row"], s.d_mat[s.sep])
    np.testing.assert_allclose(golden["projvols_cm3"], s.projvols_cm3, rtol=0, atol=0)


def test_loglike_matches_reference(golden, cl1226_oracle):
    s = cl1226_oracle
    ll = orc.get_likelihood_many(golden["thetas"], s)
    ref = golden["ll"]
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    assert np.isfinite(ref).sum() >= 20 and (~np.isfinite(ref)).sum() >= 5
    ok = np.isfinite(ref)
    # same arithmetic, same libraries: agreement to rounding
    assert np.max(np.abs(ll[ok] - ref[ok])) < 1e-9


def test_stage_outputs_match_reference(golden, cl1226_oracle):
    s = cl1226_oracle
    for w in range(0, golden["thetas"].shape[0], 3):
        p = s.full_params(golden["thetas"][w])
        with np.errstate(all="ignore"):
            st = orc.sz_stages(p, s)
            profs = np.array(orc.xray_profiles(p, s))
            mass = orc.mass_fun(p, s.r_pp, s.dens_mode)
        np.testing.assert_allclose(st["pp"], golden["pp"][w], rtol=1e-14)
        np.testing.assert_allclose(st["t_prof"], golden["tsz"][w], rtol=1e-14)
        np.testing.assert_allclose(st["bright"], golden["bright"][w], rtol=1e-11, atol=1e-14)
        assert abs(st["chisq"] - golden["chisq"][w]) <= 1e-9 * max(1.0, abs(golden["chisq"][w]))
        np.testing.assert_allclose(profs, golden["xprofs"][w], rtol=1e-13)
        np.testing.assert_allclose(mass, golden["mass"][w], rtol=1e-13)
        if profs.min() > 0:
            assert abs(orc.xray_like_from_profs(profs, s) - golden["xlike"][w]) < 1e-8


def test_batched_oracle_equals_literal(golden, cl1226_oracle):
    s = cl1226_oracle
    bo = orc.BatchedOracle(s)
    thetas = golden["thetas"][:16]
    a = bo.loglike(thetas)
    b = golden["ll"][:16]
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    ok = np.isfinite(b)
    assert np.max(np.abs(a[ok] - b[ok])) < 1e-7
    # collapsed linear operator == staged pipeline
    p = s.full_params(thetas[0])
    with np.errstate(all="ignore"):
        st = orc.sz_stages(p, s)
    row = bo.map_row(thetas[:1])[0]
    c = s.d_mat.shape[0] // 2
    assert rel_err(row, st["map_out"][c, c:]) < 1e-9
