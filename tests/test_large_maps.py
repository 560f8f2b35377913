"""Maps larger than the shipped cluster's (BASELINE configs 3 and 5: 512-point grid with a ~256-pixel map,
1024-point grid with a ~512-pixel map).  The reference can only build odd map sides (2m+1,
joxsz_main.py:101-103), so the synthetic clusters use 255 and 511 pixels; their cyclic convolution lengths
are 512 and 1024, which route the map stage to the large-map kernel (k3l2_szmap.cu; k3l_szmap.cu for wide beams).
Parity is against the oracle's literal per-walker path on seeded draws."""
import os

import numpy as np
import pytest

import kernel_model as km
from helpers import oracle_setup_from_fit, orc, rel_err_max

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _build(map_half, nr, **overrides):
    from joxsz_b200 import cluster
    from joxsz_b200.mb import mb
    mb.fit.debugfit = False
    base = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    inp = cluster.synthetic_inputs(map_half=map_half, nr=nr, base=base, **overrides)
    fit, _ = cluster.build_fit(inp, savedir=None)
    return fit


@pytest.fixture(scope="module")
def fit255():
    return _build(127, 512)


@pytest.fixture(scope="module")
def fit511():
    return _build(255, 1024)


# two more sizes on the shared-memory map kernel (cyclic length 256): a smaller map (H = 71: direct y convolution,
# filter GEMM with 9 column tiles) and a larger one (H = 101 > 88: the FFT form of the y convolution at 384
# threads, filter GEMM with 13 column tiles)
@pytest.fixture(scope="module")
def fit141():
    return _build(70, 320)


@pytest.fixture(scope="module")
def fit201():
    return _build(100, 320)


# an odd quarter plane that is not a multiple of the large-map kernel's 16-row blocks (H = 121, cyclic length 512): a
# last row pair with one row, a partial last block, convolution inputs that run past the zero rows of a tile
@pytest.fixture(scope="module")
def fit241():
    return _build(120, 384)


# a beam wider than 55 pixels (FWHM 26"): no direct y convolution, the L2-staged kernel convolves through column FFTs;
# the quarter plane (H > 136) also takes two column blocks in the filter GEMM
@pytest.fixture(scope="module")
def fitwide():
    return _build(127, 512, beam_fwhm=26.0)


def _draws(fit, n, seed, frac_bad=0.1):
    from joxsz_b200.synthetic import draw_parameters
    return draw_parameters(fit.thawed, n=n, seed=seed, spread=0.03, frac_bad=frac_bad)


@pytest.mark.parametrize("which,P", [("fit141", 256), ("fit201", 256), ("fit241", 512), ("fit255", 512), ("fit511", 1024), ("fitwide", 512)])
def test_tables_and_kernel_algorithm_on_cpu(which, P, request):
    """The packed tables + the kernel's sequence of operations (numpy model) reproduce the oracle's filtered row."""
    from joxsz_b200.packer import PackedSetup
    fit = request.getfixturevalue(which)
    pk = PackedSetup(fit, max_walkers=8)
    assert pk.map_ops.P == P and pk.N % 2 == 1
    assert np.max(np.abs(pk.bhat - pk.bhat.T)) <= 1e-14 * np.max(np.abs(pk.bhat))    # the kernel reads rows for columns
    s = oracle_setup_from_fit(fit)
    th = _draws(fit, 2, 3, frac_bad=0.0)
    pp = np.array([orc.press_fun(s.full_params(t), s.r_pp) for t in th])
    st = km.model_stages(pk, pp)
    c = pk.N // 2
    for k in range(2):
        with np.errstate(all="ignore"):
            ref = orc.sz_stages(s.full_params(th[k]), s)
        assert rel_err_max(st["row"][k], ref["map_out"][c, c:]) < 1e-12
        assert rel_err_max(st["conv"][k], ref["conv_2d"][c:, c:]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("which,nw", [("fit141", 48), ("fit201", 48), ("fit241", 48), ("fit255", 48), ("fit511", 16), ("fitwide", 32)])
def test_large_map_loglike_matches_oracle(which, nw, request):
    from joxsz_b200.batched import BatchedLikelihood
    fit = request.getfixturevalue(which)
    s = oracle_setup_from_fit(fit)
    eng = BatchedLikelihood(fit, max_walkers=512)
    th = _draws(fit, nw, 11)
    ll = eng(th)
    ref = orc.get_likelihood_many(th, s)
    assert not np.isnan(ll).any()
    assert np.array_equal(np.isfinite(ll), np.isfinite(ref))
    ok = np.isfinite(ref)
    assert ok.sum() >= nw // 2
    assert np.max(np.abs(ll[ok] - ref[ok])) < 1e-6, np.max(np.abs(ll[ok] - ref[ok]))
    # stage taps on two walkers
    good = th[ok][:2]
    maps = eng.sz_maps(good)
    prof = eng.sz_profile(good)
    c = eng.packed.N // 2
    for k in range(2):
        with np.errstate(all="ignore"):
            st = orc.sz_stages(s.full_params(good[k]), s)
        for name in ("y_2d", "conv_2d", "map_out"):
            err = rel_err_max(maps[name][k], st[name])
            assert err < 1e-5 and err < 1e-9, (name, err)
        assert rel_err_max(prof["row"][k], st["map_out"][c, c:]) < 1e-9
        assert rel_err_max(prof["bright"][k], st["bright"]) < 1e-9
        assert abs(prof["chisq"][k] - st["chisq"]) < 1e-6
    eng.close()


@pytest.mark.gpu
def test_large_map_batch_properties(fit255):
    """BASELINE config 3 scale (8192 walkers): determinism, permutation equivariance, batch independence."""
    from joxsz_b200.batched import BatchedLikelihood
    eng = BatchedLikelihood(fit255, max_walkers=8192)
    th = _draws(fit255, 8192, 21)
    ll = eng(th)
    assert not np.isnan(ll).any() and np.isfinite(ll).sum() > 4000
    perm = np.random.default_rng(0).permutation(8192)
    assert np.array_equal(eng(th[perm]), ll[perm])
    assert np.array_equal(eng(th[:777]), ll[:777])
    eng.close()
