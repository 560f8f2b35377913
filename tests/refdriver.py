"""Load the UNMODIFIED reference driver ``joxsz_main.py`` with this package standing in for its imports.

What is substituted in ``sys.modules`` (and restored afterwards):

* ``joxsz_funcs``  -> the repo-root drop-in module (this package's readers, components, likelihood, ``mcmc_run``)
* ``mbproj2``      -> ``joxsz_b200.mb.mb`` (the real mbproj2 when importable, else the bundled work-alike)
* ``emcee``        -> ``EnsembleSampler`` = :class:`joxsz_b200.sampler.EnsembleSampler`; ``backends.HDFBackend`` raises
                      (h5py is absent; the reference's own ``try/except`` at ``joxsz_main.py:197-201`` then runs without it)
* ``joxsz_plots``  -> the batched posterior computations of :mod:`joxsz_b200.posterior` under the reference's names,
                      the matplotlib/corner plotting functions as no-ops (plots are out of scope)

XSPEC cannot build the count-rate tables, so ``CountRate.addCountCache`` fills the synthetic tables of
``joxsz_b200.synthetic`` for whatever key it is asked for -- the hook the reference itself leaves at
``joxsz_main.py:189``.
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REFERENCE_DIR = "/root/reference"


def reference_main_available():
    return os.path.exists(os.path.join(REFERENCE_DIR, "joxsz_main.py")) and os.path.isdir(
        os.path.join(REFERENCE_DIR, "data"))


class _Swap:
    """Context manager: install module stubs / attribute patches and undo them on exit."""

    def __init__(self):
        self._mods = {}
        self._attrs = []

    def module(self, name, mod):
        self._mods[name] = sys.modules.get(name)
        sys.modules[name] = mod

    _MISSING = object()

    def guard(self, obj, name):
        """Remember the current state of ``obj.name`` (an own attribute, or absent) for :meth:`restore`."""
        self._attrs.append((obj, name, obj.__dict__.get(name, self._MISSING), self._MISSING))

    def attr(self, obj, name, value):
        self.guard(obj, name)
        setattr(obj, name, value)

    def restore(self):
        for obj, name, old, missing in reversed(self._attrs):
            if old is missing:
                try:
                    delattr(obj, name)
                except AttributeError:
                    pass
            else:
                setattr(obj, name, old)
        for name, old in self._mods.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


def load_reference_main(outdir, nburn=4, nlength=6, nthin=2, nwalkers=30, seed=11, max_prefit=None):
    """Returns ``(module, swap)``: the reference's ``joxsz_main`` module ready for ``module.main()`` (call with the
    working directory at ``REFERENCE_DIR``: its data paths are relative), and the swap object to ``restore()``."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from joxsz_b200 import posterior, sampler
    from joxsz_b200.mb import mb
    from joxsz_b200.synthetic import synthetic_countrate_tables
    import joxsz_funcs as dropin

    sw = _Swap()
    sw.module("mbproj2", mb)
    sw.module("joxsz_funcs", dropin)

    emcee = types.ModuleType("emcee")
    backends = types.ModuleType("emcee.backends")

    def HDFBackend(*a, **k):
        raise ImportError("h5py is not installed: chains are kept in memory / written as .npz")

    backends.HDFBackend = HDFBackend
    emcee.backends = backends
    emcee.EnsembleSampler = sampler.EnsembleSampler
    sw.module("emcee", emcee)
    sw.module("emcee.backends", backends)

    plots = types.ModuleType("joxsz_plots")
    for nm in ("best_fit_prof", "comp_rad_profs", "comp_mass_prof", "frac_gas_prof"):
        setattr(plots, nm, getattr(posterior, nm))
    calls = []
    for nm in ("traceplot", "triangle", "fitwithmod", "plot_rad_profs", "mass_plot", "frac_gas_plot"):
        setattr(plots, nm, (lambda n: (lambda *a, **k: calls.append(n)))(nm))
    plots.calls = calls
    sw.module("joxsz_plots", plots)

    def addCountCache(self, key):
        emin, emax = key[0], key[1]
        t0, t1 = synthetic_countrate_tables([(emin, emax)], self.Tlogvals)[0]
        self.ctcache[key] = (np.asarray(t0, dtype=np.float64), np.asarray(t1, dtype=np.float64))

    sw.attr(mb.countrate.CountRate, "addCountCache", addCountCache)
    # main() rebinds these class attributes (joxsz_main.py:129-137, 186-188): put them back afterwards
    for cls, names in ((mb.Fit, ("get_sz_like", "getLikelihood", "mylikeFromProfs")),
                       (mb.CmptFlat, ("defPars",)), (mb.CmptVikhDensity, ("vikhFunction", "defPars", "prior"))):
        for nm in names:
            sw.guard(cls, nm)

    if max_prefit is not None:
        # the reference's preliminary loop runs until the best log-probability stops improving; bound it for tests
        def mcmc_run(mcmc, fit, nburn, nsteps, nthin=1, **kw):
            return sampler.mcmc_run(mcmc, fit, nburn, nsteps, nthin, max_prefit=max_prefit, **kw)
        sw.attr(dropin, "mcmc_run", mcmc_run)

    spec = importlib.util.spec_from_file_location("ref_joxsz_main", os.path.join(REFERENCE_DIR, "joxsz_main.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sw.restore()
        raise
    out = str(outdir).rstrip("/") + "/"
    mod.savedir = mod.plotdir = out
    mod.nburn, mod.nlength, mod.nthin, mod.nwalkers, mod.seed = nburn, nlength, nthin, nwalkers, seed
    mod._plot_calls = calls
    return mod, sw
