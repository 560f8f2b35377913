"""The reference's own driver against this package (drop-in boundary).

* CPU, where ``/root/reference`` exists: ``joxsz_main.main()`` is imported unmodified (tests/refdriver.py) and run up
  to the first likelihood evaluation -- every reader, component, parameter and binding of ``joxsz_main.py:93-188`` goes
  through this package; the first ``getLikelihood`` then fails loudly because there is no GPU (no CPU fallback).
* CPU: a fit with a device engine attached must pickle (``joxsz_main.py:193-194`` pickles it right after ``doFitting``).
* GPU: the sequence of ``joxsz_main.py:186-215`` (class-level binding, doFitting, pickle, sampler with pool/backend
  arguments, mcmc_run, chain read-out) on the committed input fixture; and, where the reference tree is present next to
  a GPU, ``main()`` itself from start to end.
"""
import os
import pickle
from types import MethodType

import numpy as np
import pytest

import refdriver

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(not refdriver.reference_main_available(), reason="needs /root/reference (build container only)")
def test_reference_main_setup_runs_through_this_package(tmp_path, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU test below")
    from joxsz_b200 import _lib
    monkeypatch.chdir(refdriver.REFERENCE_DIR)
    mod, sw = refdriver.load_reference_main(tmp_path)
    try:
        # everything up to joxsz_main.py:190 is host set-up; :191 (fit.doFitting) evaluates the likelihood -> needs CUDA
        with pytest.raises(_lib.JxError, match="CUDA"):
            mod.main()
        from joxsz_b200.mb import mb
        bound = mb.Fit.__dict__["getLikelihood"]
        fit = bound.__self__
        assert isinstance(bound, MethodType) and bound.__func__.__module__ == "joxsz_b200.funcs"
        assert len(fit.thawed) == 13 and fit.data.sz.r_pp.size == 313 and fit.data.sz.d_mat.shape == (171, 171)
        assert len(fit.data.bands) == 10 and fit.data.annuli.nshells == 15
        # the set-up main() built is the one the committed fixture rebuilds (same readers, same order)
        from joxsz_b200 import cluster
        inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
        fit2, _ = cluster.build_fit(inp, savedir=None)
        assert list(fit2.thawed) == list(fit.thawed)
        for k in ("beam_2d", "filtering", "r_pp", "radius", "flux_data"):
            np.testing.assert_array_equal(np.asarray(getattr(fit.data.sz, k)), np.asarray(getattr(fit2.data.sz, k)))
        pickle.dumps(fit, -1)
    finally:
        sw.restore()


class _FakeEngine:
    """Stands in for a BatchedLikelihood: holds what makes the real one unpicklable (a ctypes pointer)."""
    max_walkers = 1 << 20

    def __init__(self):
        import ctypes
        self._h = ctypes.c_void_p(1234)
        self.closed = False

    def close(self):
        self.closed = True


def test_fit_with_engine_attached_pickles(cl1226_fit):
    """joxsz_main.py:193-194: ``pickle.dump(fit, f, -1)`` after the engine exists; emcee pickles the bound method."""
    import weakref
    from joxsz_b200 import funcs
    fit = cl1226_fit
    eng = _FakeEngine()
    funcs._ENGINES[id(fit)] = (weakref.ref(fit), eng, funcs._signature(fit))
    try:
        assert funcs.engine_for(fit, 16) is eng                 # the cache is what getLikelihood would use
        assert not any(k.startswith("_jx") for k in fit.__dict__)
        blob = pickle.dumps(fit, -1)
        fit2 = pickle.loads(blob)
        assert list(fit2.thawed) == list(fit.thawed)
        assert id(fit2) not in funcs._ENGINES                   # a copy gets its own engine on first use
        pickle.dumps(fit.getLikelihood, -1)
    finally:
        funcs.jx_invalidate(fit)
    assert eng.closed and id(fit) not in funcs._ENGINES


def test_engine_cache_follows_the_setup(cl1226_fit):
    """A changed frozen value, N_H or data array must not keep serving the old device engine."""
    from joxsz_b200 import funcs
    fit = cl1226_fit
    s0 = funcs._signature(fit)
    nh = fit.model.NH_1022pcm2
    fit.model.NH_1022pcm2 = nh * 2
    assert funcs._signature(fit) != s0
    fit.model.NH_1022pcm2 = nh
    assert funcs._signature(fit) == s0
    old = fit.data.sz.flux_data
    fit.data.sz.flux_data = np.array(old, copy=True)
    assert funcs._signature(fit) != s0
    fit.data.sz.flux_data = old
    assert funcs._signature(fit) == s0


def _driver_tail(fit, mb, mc, mcmc_run, add_backend_attrs, savedir, nburn, nlength, nwalkers, nthin, seed, name="joxsz"):
    """joxsz_main.py:186-215 with the module globals as arguments."""
    from multiprocessing import Pool
    from joxsz_b200 import funcs
    mb.Fit.get_sz_like = MethodType(funcs.get_sz_like, fit)
    mb.Fit.getLikelihood = MethodType(funcs.getLikelihood, fit)
    mb.Fit.mylikeFromProfs = MethodType(funcs.mylikeFromProfs, fit)
    fit.doFitting()
    with open('%s%s_fit.pickle' % (savedir, name), 'wb') as f:
        pickle.dump(fit, f, -1)
    chainfilename = '%s%s_chain.hdf5' % (savedir, name)
    backend = None
    try:
        backend = mc.backends.HDFBackend(chainfilename)
        backend.reset(nwalkers, len(fit.thawedParVals()))
    except:  # noqa: E722  (the reference's bare except)
        pass
    with Pool(2) as pool:
        np.random.seed(seed)
        try:
            mcmc = mc.EnsembleSampler(nwalkers, len(fit.thawed), fit.getLikelihood, pool=pool, backend=backend)
        except:  # noqa: E722
            mcmc = mc.EnsembleSampler(nwalkers, len(fit.thawed), fit.getLikelihood, pool=pool)
        mcmc.initspread = .1
        mcmc_run(mcmc, fit, nburn, nlength, nthin)
        try:
            add_backend_attrs(chainfilename, fit, nburn, nthin)
        except (ImportError, OSError):
            pass            # no h5py here: the chain stays in memory (save_npz is the offered format)
    return mcmc


@pytest.mark.gpu
def test_driver_sequence_after_setup(tmp_path):
    """joxsz_main.py:186-215 on the committed fixture: class-level binding, doFitting, pickle of the fit with its
    engine alive, EnsembleSampler(..., pool=, backend=), mcmc_run, chain read-out."""
    import types
    from joxsz_b200 import cluster, funcs, sampler
    from joxsz_b200.mb import mb
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    fit, _ = cluster.build_fit(inp, savedir=str(tmp_path))
    for nm in ("get_sz_like", "getLikelihood", "mylikeFromProfs"):     # main() binds on the class, not the instance
        fit.__dict__.pop(nm, None)
    mc = types.SimpleNamespace(EnsembleSampler=sampler.EnsembleSampler,
                               backends=types.SimpleNamespace(HDFBackend=lambda *a, **k: (_ for _ in ()).throw(ImportError())))
    saved = {nm: mb.Fit.__dict__.get(nm) for nm in ("get_sz_like", "getLikelihood", "mylikeFromProfs")}

    def run(mcmc, fit, nburn, nsteps, nthin):
        return sampler.mcmc_run(mcmc, fit, nburn, nsteps, nthin, max_prefit=1)
    try:
        mcmc = _driver_tail(fit, mb, mc, run, funcs.add_backend_attrs, str(tmp_path) + "/", nburn=6, nlength=10,
                            nwalkers=30, nthin=5, seed=3)
        assert id(fit) in funcs._ENGINES                               # the engine existed when the fit was pickled
        fit2 = pickle.load(open(str(tmp_path) + "/joxsz_fit.pickle", "rb"))
        assert list(fit2.thawed) == list(fit.thawed)
        cube = mcmc.chain                                              # joxsz_main.py:213
        assert cube.shape == (30, 2, 13) and np.all(np.isfinite(cube))
        flat = cube.reshape(-1, cube.shape[2], order='F')
        assert np.all(np.isfinite(np.median(flat, axis=0)))
        assert np.all(np.isfinite(mcmc.get_log_prob()))
    finally:
        for nm, v in saved.items():
            if v is None:
                if nm in mb.Fit.__dict__:
                    delattr(mb.Fit, nm)
            else:
                setattr(mb.Fit, nm, v)
        funcs.jx_invalidate(fit)


@pytest.mark.gpu
@pytest.mark.skipif(not refdriver.reference_main_available(), reason="needs /root/reference next to a GPU")
def test_reference_main_end_to_end(tmp_path, monkeypatch):
    """The unmodified ``main()`` from its first line to its last (plots are no-ops, posterior summaries are real)."""
    monkeypatch.chdir(refdriver.REFERENCE_DIR)
    mod, sw = refdriver.load_reference_main(tmp_path, max_prefit=1)
    try:
        mod.main()
        assert os.path.exists(os.path.join(str(tmp_path), "joxsz_fit.pickle"))
        assert mod._plot_calls == ["traceplot", "triangle", "fitwithmod", "plot_rad_profs", "mass_plot", "frac_gas_plot"]
    finally:
        sw.restore()
