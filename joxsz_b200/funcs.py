"""The likelihood methods JoXSZ binds onto ``mbproj2.Fit`` -- same names, arguments and return
conventions as ``joxsz_funcs.py:439-546`` -- routed to the batched CUDA engine.

``getLikelihood(self, vals=None)`` accepts what the reference accepts (``None`` or one thawed vector)
and additionally ``[W, ndim]`` arrays (numpy or CUDA torch float64), returning a float or ``[W]``.
``get_sz_like(self, output=...)`` serves 'll', 'chisq', 'pp', 'bright' from the parity taps.  There is no
host implementation behind these: without the CUDA library / a GPU they raise.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .mb import mb


# ----------------------------------------------------------------------------------------------
# engine cache: one BatchedLikelihood per fit object, rebuilt when the frozen set-up changes
# ----------------------------------------------------------------------------------------------

def _arr_id(a):
    """Cheap identity of a data array: a replaced or reshaped array changes it (in-place edits do not: call
    :func:`jx_invalidate` after those)."""
    if a is None:
        return None
    try:
        return (id(a), tuple(np.shape(a)))
    except Exception:
        return id(a)


def _signature(fit):
    """Everything the device engine bakes in at ``jx_create``: thawed names, priors, frozen values, N_H, the density
    mode and the identity of every data array (SZ data, beam, filter, grids, bands, count-rate tables)."""
    sz = fit.data.sz
    sig = [tuple(fit.thawed), bool(getattr(fit, "exclude_unphy_mass", False)),
           (bool(getattr(sz, "calc_integ", False)), getattr(sz, "integ_mu", None), getattr(sz, "integ_sig", None))]
    for name, par in fit.pars.items():
        if name in fit.thawed:
            if hasattr(par, "prior_mu"):
                sig.append((name, "g", float(par.prior_mu), float(par.prior_sigma)))
            else:
                sig.append((name, "b", float(par.minval), float(par.maxval)))
        else:
            sig.append((name, "f", float(np.asarray(par.val).reshape(-1)[0])))
    model = getattr(fit, "model", None)
    sig.append(("NH", float(getattr(model, "NH_1022pcm2", 0.0) or 0.0)))
    sig.append(("dens", getattr(getattr(model, "ne_cmpt", None), "mode", None)))
    for k in ("flux_data", "beam_2d", "filtering", "r_pp", "radius", "d_mat", "convert"):
        sig.append((k, _arr_id(getattr(sz, k, None))))
    sig.append(("step", getattr(sz, "step", None), getattr(sz, "kpc_as", None), getattr(sz, "sep", None)))
    bands = getattr(fit.data, "bands", None) or ()
    for b in bands:
        sig.append(("band", _arr_id(getattr(b, "cts", None)), _arr_id(getattr(b, "exposures", None)),
                    _arr_id(getattr(b, "backrates", None)), _arr_id(getattr(b, "areascales", None))))
    ctr = getattr(getattr(fit.data, "annuli", None), "ctrate", None)
    cache = getattr(ctr, "ctcache", None)
    if cache is not None:
        sig.append(("ctcache", id(cache), len(cache)))
    return tuple(sig)


# The engine holds ctypes pointers and device memory: it must not live in ``fit.__dict__`` -- the reference pickles the
# fit right after ``doFitting`` (``joxsz_main.py:193-194``) and emcee pickles ``fit.getLikelihood`` for its pool.
# Engines are kept here, keyed by the identity of the fit object and dropped when the fit is collected.
_ENGINES = {}


def _drop_engine(key):
    ent = _ENGINES.pop(key, None)
    if ent is not None:
        try:
            ent[1].close()
        except Exception:
            pass


def jx_invalidate(fit):
    """Forget the device engine of ``fit`` (call after editing one of its data arrays in place)."""
    _drop_engine(id(fit))


def attach_engine(fit, eng):
    """Make ``eng`` (a :class:`BatchedLikelihood` built for ``fit``) the engine ``fit.getLikelihood`` uses, instead of
    letting the first call build one (a caller that already sized an engine for its ensemble)."""
    import weakref
    key = id(fit)
    ent = _ENGINES.get(key)
    if ent is not None and ent[1] is not eng:
        _drop_engine(key)
    _ENGINES[key] = (weakref.ref(fit, lambda _r, k=key: _drop_engine(k)), eng, _signature(fit))


def detach_engine(fit):
    """Forget the engine of ``fit`` without closing it (the caller owns it)."""
    _ENGINES.pop(id(fit), None)


def engine_for(fit, min_walkers=1):
    """The fit's :class:`BatchedLikelihood`, (re)built on demand."""
    import weakref
    from .batched import BatchedLikelihood
    key = id(fit)
    ent = _ENGINES.get(key)
    if ent is not None and ent[0]() is not fit:          # a dead object's id was reused
        _drop_engine(key)
        ent = None
    sig = _signature(fit)
    if ent is not None and ent[2] == sig and ent[1].max_walkers >= min_walkers:
        return ent[1]
    if ent is not None:
        _drop_engine(key)
    cap = max(int(min_walkers), int(getattr(fit, "jx_max_walkers", 1024)))
    eng = BatchedLikelihood(fit, max_walkers=cap)
    try:
        ref = weakref.ref(fit, lambda _r, k=key: _drop_engine(k))
    except TypeError:                                    # not weak-referenceable: keep it alive with the engine
        ref = (lambda f: (lambda: f))(fit)
    _ENGINES[key] = (ref, eng, sig)
    return eng


def _current_theta(fit):
    """Thawed values as [W, ndim] (W = 1 unless some ``.val`` is walker-valued)."""
    cols = [np.asarray(fit.pars[n].val, dtype=np.float64).reshape(-1) for n in fit.thawed]
    W = max(c.size for c in cols)
    return np.stack([np.broadcast_to(c, (W,)) for c in cols], axis=1), W


def _squeeze(a, W_is_one):
    return a[0] if W_is_one else a


# ----------------------------------------------------------------------------------------------
# methods bound onto mb.Fit (joxsz_main.py:186-188)
# ----------------------------------------------------------------------------------------------

def get_sz_like(self, output="ll"):
    """SZ log-likelihood (or an intermediate) for the current parameters; reference ``joxsz_funcs.py:439-493``.

    output: 'll' | 'chisq' | 'pp' (pressure profile on r_pp) | 'bright' (surface-brightness profile) |
    'integ' (integrated Compton parameter; only with ``calc_integ=True``, like the reference).
    """
    theta, W = _current_theta(self)
    eng = engine_for(self, W)
    one = W == 1
    if output == "pp":
        return _squeeze(eng.profiles(theta)["pp"], one)
    if output in ("bright", "ll", "chisq", "integ"):
        res = eng.sz_profile(theta)
        if output == "bright":
            return _squeeze(res["bright"], one)
        chisq = res["chisq"]
        calc_integ = bool(getattr(self.data.sz, "calc_integ", False))
        if output == "integ":
            if not calc_integ:      # the reference falls through to its RuntimeError when calc_integ is off
                raise RuntimeError('Unrecognised output name (must be "ll", "chisq", "pp", "bright" or "integ")')
            val = res["cint"]
        elif output == "ll":
            val = -chisq / 2
            if calc_integ:
                with np.errstate(invalid="ignore"):
                    pen = ((res["cint"] - self.data.sz.integ_mu) / self.data.sz.integ_sig) ** 2
                val = val - np.where(np.isnan(pen), 0.0, pen) / 2
        else:
            val = chisq
        return float(val[0]) if one else val
    raise RuntimeError('Unrecognised output name (must be "ll", "chisq", "pp", "bright" or "integ")')


def calcProfiles(self):
    """Predicted X-ray count profiles per band for the current parameters (mbproj2 ``Fit.calcProfiles``,
    called at ``joxsz_funcs.py:527``): list of nb arrays [na] (or [W, na])."""
    theta, W = _current_theta(self)
    pred = engine_for(self, W).xray(theta)["pred"]
    return [pred[0, b] if W == 1 else pred[:, b] for b in range(pred.shape[1])]


def mylikeFromProfs(self, predprofs):
    """Cash log-likelihood of given predicted profiles over bins with non-NaN counts
    (reference ``joxsz_funcs.py:495-505``).  Accepts nb arrays [na] or [W, na]."""
    pred = np.stack([np.asarray(p, dtype=np.float64) for p in predprofs], axis=-2)   # [..., nb, na]
    one = pred.ndim == 2
    pred = pred.reshape((-1,) + pred.shape[-2:])
    out = engine_for(self, pred.shape[0]).cash_from_profiles(pred)
    return float(out[0]) if one else out


def getLikelihood(self, vals=None):
    """Joint X-SZ log-likelihood; reference ``joxsz_funcs.py:507-546``.

    ``vals``: None (current parameters), one thawed vector, or [W, ndim].  Returns a float, or [W]
    (numpy, or a CUDA tensor if a CUDA tensor was given).  -inf exactly where the reference returns
    -inf.  Side effects kept: parameters are updated (single vector: to ``vals``; batch: to the best
    walker when it improves ``bestlike`` by more than 0.1) and ``fit.dat`` is rewritten on improvement.
    """
    import torch
    if vals is None:
        theta, W = _current_theta(self)
        single = W == 1
    elif isinstance(vals, torch.Tensor):
        theta = vals
        single = vals.dim() == 1
        W = 1 if single else vals.shape[0]
    else:
        theta = np.asarray(vals, dtype=np.float64)
        single = theta.ndim == 1
        W = 1 if single else theta.shape[0]
        if single:
            self.updateThawed(theta)
    eng = engine_for(self, W)
    if not isinstance(theta, torch.Tensor):
        theta = theta.reshape(-1, theta.shape[-1])       # [W, ndim]; W = 1 for a single vector
    ll = eng(theta)
    if isinstance(ll, torch.Tensor) and ll.is_cuda:
        best_val, best_idx = (ll, 0) if ll.dim() == 0 else torch.max(ll, dim=0)
        best_val, best_idx = float(best_val), int(best_idx)
    else:
        arr = np.atleast_1d(np.asarray(ll, dtype=np.float64))
        best_idx = int(np.argmax(arr))
        best_val = float(arr[best_idx])
    if mb.fit.debugfit and (best_val - self.bestlike) > 0.1:
        best_theta = theta[best_idx] if theta.ndim == 2 else theta
        if isinstance(best_theta, torch.Tensor):
            best_theta = best_theta.detach().cpu().numpy()
        best_theta = np.asarray(best_theta, dtype=np.float64).reshape(-1)
        self.updateThawed(best_theta)
        self.bestlike = best_val
        _write_fit_dat(self, eng, best_theta, best_val)
    if single:
        if isinstance(ll, torch.Tensor):
            return float(ll.reshape(-1)[0])
        return float(ll[0]) if np.ndim(ll) else float(ll)
    return ll


def _write_fit_dat(fit, eng, theta, totlike):
    """``fit.dat`` dump of reference ``joxsz_funcs.py:540-545`` (same text layout)."""
    savedir = getattr(fit, "savedir", None)
    if savedir is None:
        return
    th = theta[None, :]
    like = float(eng.xray(th)["cash"][0])
    sz_like = float(-eng.sz_profile(th)["chisq"][0] / 2)
    prior = totlike - like - sz_like
    try:
        with mb.utils.AtomicWriteFile("%s/fit.dat" % savedir) as fout:
            mb.utils.uprint("likelihood = %g + %g + %g = %g" % (like, sz_like, prior, totlike), file=fout)
            for p in sorted(fit.pars):
                mb.utils.uprint("%s = %s" % (p, fit.pars[p]), file=fout)
    except OSError:
        pass


# ----------------------------------------------------------------------------------------------
# remaining names joxsz_main.py imports
# ----------------------------------------------------------------------------------------------

def addCountCache(self, key):
    """Count-rate table builder (reference ``joxsz_funcs.py:652-681``) -- needs XSPEC, which this
    package does not drive.  Tables are supplied through ``CountRate.ctcache`` instead."""
    raise RuntimeError("XSPEC-backed count-rate tables are out of scope here: fill "
                       "annuli.ctrate.ctcache[key] = (ln rate_Z0, ln rate_Z1) yourself "
                       "(see joxsz_b200.synthetic.synthetic_countrate_tables for the format)")


def add_backend_attrs(chainfilename, fit, nburn, nthin):
    """Attach ``param_names`` / ``burn`` / ``thin`` to a saved chain (reference ``joxsz_funcs.py:637-650``).
    HDF5 chains need h5py; ``.npz`` chains written by :mod:`joxsz_b200.sampler` are updated in place."""
    if str(chainfilename).endswith(".npz"):
        z = dict(np.load(chainfilename, allow_pickle=False))
        z["param_names"] = np.array([k.encode("utf-8") for k in fit.thawed])
        z["burn"], z["thin"] = np.array(nburn), np.array(nthin)
        np.savez_compressed(chainfilename, **z)
        return
    import h5py  # noqa: raises ImportError where unavailable
    with h5py.File(chainfilename, "r+") as f:
        f["mcmc"].attrs["param_names"] = np.array([k.encode("utf-8") for k in fit.thawed])
        f["mcmc"].attrs["burn"] = nburn
        f["mcmc"].attrs["thin"] = nthin
