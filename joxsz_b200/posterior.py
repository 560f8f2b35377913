"""Batched posterior post-processing: what ``joxsz_plots.py`` computes sample by sample from a chain,
evaluated for all selected samples at once on the GPU.

Same function names, arguments and return shapes as the reference (``joxsz_plots.py``):
``get_equal_tailed`` (:93-101), ``best_fit_prof`` (:104-132), ``frac_int`` (:194-206), ``cum_gas_mass``
(:208-217), ``thermodynamic_profs`` (:219-247), ``comp_rad_profs`` (:249-273), ``hydro_mass`` (:316-339),
``comp_mass_prof`` (:341-376), ``mass_overdens`` (:378-399), ``frac_gas_prof`` (:451-476).  The reference
loops ``fit.updateThawed(sample)`` + per-sample calls; here the per-sample model evaluations (X-ray
profiles, SZ brightness, density / pressure / temperature / hydrostatic mass on a radial grid) run in
batches through the C-ABI taps, and only the reductions that are pure bookkeeping (percentiles, the
cumulative-sum gas mass, the critical-density formula) stay in numpy.  Plotting itself is out of scope.

Sample selection reproduces the reference: ``np.random.seed(seed); np.random.choice(nw*niter, num,
replace=False)`` over the ``meshgrid`` ordering of (walker, iteration).
"""
from __future__ import annotations

import numpy as np

from . import radial
from .funcs import engine_for
from .mb import mb

_pc = mb.physconstants


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------

def get_equal_tailed(data, ci=95):
    """Median and equal-tailed interval along axis 0 -> array [3, ...] (``joxsz_plots.py:93-101``)."""
    low, med, upp = map(np.atleast_1d, np.percentile(data, [50 - ci / 2, 50, 50 + ci / 2], axis=0))
    return np.array([low, med, upp])


def select_samples(cube_chain, num="all", seed=None):
    """The ``[num, ndim]`` parameter sets the reference's loops visit, in its order."""
    cube_chain = np.asarray(cube_chain)
    nw, nit = cube_chain.shape[0], cube_chain.shape[1]
    if num == "all":
        num = nw * nit
    w, it = np.meshgrid(np.arange(nw), np.arange(nit))
    w, it = w.flatten(), it.flatten()
    np.random.seed(seed)
    rand = np.random.choice(w.size, num, replace=False)
    return np.ascontiguousarray(cube_chain[w[rand], it[rand], :], dtype=np.float64)


class _Val:
    __slots__ = ("val",)

    def __init__(self, val):
        self.val = val


def batched_pars(fit, samples):
    """``fit.pars``-like mapping whose thawed entries carry one value per sample (the component methods
    accept walker-valued ``.val``), frozen entries their scalar."""
    samples = np.asarray(samples, dtype=np.float64)
    out = {}
    for name, par in fit.pars.items():
        out[name] = _Val(samples[:, fit.thawed.index(name)].copy() if name in fit.thawed
                         else float(np.asarray(par.val).reshape(-1)[0]))
    return out


def _batches(n, size):
    for lo in range(0, n, size):
        yield lo, min(n, lo + size)


# ----------------------------------------------------------------------------------------------
# surface-brightness profiles (joxsz_plots.py:104-132)
# ----------------------------------------------------------------------------------------------

def model_profiles(samples, fit, batch=8192):
    """X-ray predicted profiles ``[S, nb, na]`` and SZ brightness profiles ``[S, sep+1]`` of every sample
    (``fit.calcProfiles()`` and ``fit.get_sz_like(output='bright')`` of the reference's loop)."""
    samples = np.asarray(samples, dtype=np.float64)
    eng = engine_for(fit, min(batch, max(samples.shape[0], 1)))
    step = min(batch, eng.max_walkers)
    px, ps = [], []
    for lo, hi in _batches(samples.shape[0], step):
        px.append(eng.xray(samples[lo:hi])["pred"])
        ps.append(eng.sz_profile(samples[lo:hi])["bright"])
    return np.concatenate(px), np.concatenate(ps)


def best_fit_prof(cube_chain, fit, num="all", seed=None, ci=95):
    """Median and interval of the surface-brightness profiles -> ``(perc_x [3, nb, na], perc_sz [3, sep+1])``."""
    samples = select_samples(cube_chain, num, seed)
    profs_x, profs_sz = model_profiles(samples, fit)
    return get_equal_tailed(profs_x, ci), get_equal_tailed(profs_sz, ci)


# ----------------------------------------------------------------------------------------------
# thermodynamic profiles (joxsz_plots.py:194-273)
# ----------------------------------------------------------------------------------------------

def frac_int(edges):
    """Fraction of a shell's mass inside its mid-point (``joxsz_plots.py:194-206``)."""
    low_r, hig_r = edges[:-1], edges[1:]
    volinside = (low_r + hig_r) ** 3 / 24 - low_r ** 3 / 3
    voloutside = hig_r ** 3 / 3 - (low_r + hig_r) ** 3 / 24
    return volinside / (volinside + voloutside)


def cum_gas_mass(r_kpc, dens):
    """Cumulative gas mass (``joxsz_plots.py:208-217``); ``dens`` may be ``[n]`` or ``[S, n]``."""
    r_kpc = np.asarray(r_kpc, dtype=np.float64)
    dens = np.asarray(dens, dtype=np.float64)
    edg_cm = np.append(r_kpc[0] / 2, r_kpc + r_kpc[0] / 2) * _pc.kpc_cm
    mgas = dens * _pc.mu_e * _pc.mu_g / _pc.solar_mass_g * 4 / 3 * np.pi * (edg_cm[1:] ** 3 - edg_cm[:-1] ** 3)
    csum = np.cumsum(mgas, axis=-1)
    prev = np.concatenate((np.zeros(mgas.shape[:-1] + (1,)), csum[..., :-1]), axis=-1)
    return mgas * frac_int(edg_cm) + prev


def thermodynamic_profs(vals, r_kpc, fit):
    """``(dens, temp, press, entr, cool, cmgas, tempx)`` on ``r_kpc`` for one parameter vector ``[ndim]`` or a
    batch ``[S, ndim]`` (each output ``[n]`` or ``[S, n]``); reference ``joxsz_plots.py:219-247``.

    The cooling time needs ``annuli.ctrate.getFlux`` (an XSPEC flux table); without it ``cool`` is NaN.
    A single vector also updates ``fit``'s parameters like the reference does."""
    vals = np.asarray(vals, dtype=np.float64)
    single = vals.ndim == 1
    if single:
        fit.updateThawed(vals)
    pars = batched_pars(fit, np.atleast_2d(vals))
    dens = np.atleast_2d(fit.model.ne_cmpt.vikhFunction(pars, r_kpc))
    press = np.atleast_2d(fit.press.press_fun(pars, r_kpc))
    temp = press / dens
    logratio = np.asarray(pars["log(T_X/T_{SZ})"].val, dtype=np.float64).reshape(-1, 1)
    tempx = temp * 10 ** logratio
    entr = temp / dens ** (2 / 3)
    ctrate = fit.data.annuli.ctrate
    if hasattr(ctrate, "getFlux"):
        zname = getattr(fit.model.Z_cmpt, "name", "Z")
        zval = np.broadcast_to(np.asarray(pars[zname].val, dtype=np.float64).reshape(-1, 1), temp.shape)
        flux = np.stack([ctrate.getFlux(temp[i], zval[i], dens[i]) for i in range(temp.shape[0])])
        cool = (5 / 2) * dens * (1. + 1 / _pc.ne_nH) * temp * _pc.keV_erg / (
            flux * 4. * np.pi * (fit.data.annuli.cosmology.D_L * _pc.Mpc_cm) ** 2) / _pc.yr_s
    else:
        cool = np.full_like(temp, np.nan)
    cmgas = cum_gas_mass(r_kpc, dens)
    out = (dens, temp, press, entr, cool, cmgas, tempx)
    return tuple(o[0] for o in out) if single else out


def comp_rad_profs(cube_chain, fit, num="all", seed=None, ci=95, batch=16384):
    """Median and interval of the thermodynamic profiles on ``r_pp`` -> 7 arrays ``[3, Nr]``
    (``joxsz_plots.py:249-273``)."""
    samples = select_samples(cube_chain, num, seed)
    parts = [thermodynamic_profs(samples[lo:hi], fit.data.sz.r_pp, fit) for lo, hi in _batches(samples.shape[0], batch)]
    return tuple(get_equal_tailed(np.concatenate([p[k] for p in parts]), ci) for k in range(7))


# ----------------------------------------------------------------------------------------------
# hydrostatic mass, overdensity radius, gas fraction (joxsz_plots.py:316-399, 451-476)
# ----------------------------------------------------------------------------------------------

def mass_overdens(r_kpc, cosmo, delta=500):
    """Mass of a sphere of mean density ``delta * rho_crit(z)`` (``joxsz_plots.py:378-399``), solar masses."""
    H0_s = cosmo.H0 / _pc.Mpc_km
    HZ = H0_s * np.sqrt(cosmo.WM * (1. + cosmo.z) ** 3 + cosmo.WV)
    rho_c = 3. * HZ ** 2 / (8. * np.pi * _pc.G_cgs)
    r_cm = np.asarray(r_kpc, dtype=np.float64) * _pc.kpc_cm
    return 4 / 3 * np.pi * rho_c * delta * r_cm ** 3 / _pc.solar_mass_g


def _mass_at(fit, pars, r_rows):
    """Hydrostatic mass of sample s at its own radii ``r_rows[s, :]`` (one kernel launch)."""
    cm = fit.mass_cmpt
    mode = getattr(cm.ne_prof, "mode", "single")
    need = ("P_0", "a", "b", "c", "r_p", "log(n_0)", r"\beta", "log(r_c)", "log(r_s)", r"\alpha", r"\epsilon", r"\gamma")
    if mode == "double":
        need += ("log(n_{02})", r"\beta_2", "log(r_{c2})")
    return radial.evaluate(pars, r_rows, "mass", need=need, mode=mode, mu_gas=0.61, per_walker_r=True)


def overdensity_radius(samples, fit, cosmo, delta=500, start_opt=700., tol=1.48e-8, maxiter=50):
    """``r_delta`` per sample: root of ``M_HSE(r) - M_delta(r)`` by the secant iteration scipy's
    ``optimize.newton`` runs when no derivative is given (same starting pair, same tolerance), advanced for
    all samples together; samples that do not converge get NaN (scipy would raise)."""
    samples = np.atleast_2d(np.asarray(samples, dtype=np.float64))
    S = samples.shape[0]
    pars = batched_pars(fit, samples)
    f = lambda r: _mass_at(fit, pars, r.reshape(S, 1)).reshape(S) - mass_overdens(r, cosmo, delta)
    p0 = np.full(S, float(start_opt))
    eps = 1e-4
    p1 = p0 * (1 + eps) + np.where(p0 >= 0, eps, -eps)
    q0, q1 = f(p0), f(p1)
    swap = np.abs(q1) < np.abs(q0)
    p0, p1 = np.where(swap, p1, p0), np.where(swap, p0, p1)
    q0, q1 = np.where(swap, q1, q0), np.where(swap, q0, q1)
    root = np.full(S, np.nan)
    active = np.ones(S, dtype=bool)
    for _ in range(maxiter):
        with np.errstate(all="ignore"):
            denom = q1 - q0
            p = np.where(np.abs(q1) > np.abs(q0), (-q0 / q1 * p1 + p0) / (1 - q0 / q1),
                         (-q1 / q0 * p0 + p1) / (1 - q1 / q0))
        flat = active & (denom == 0)
        root[flat] = ((p1 + p0) / 2.0)[flat]
        active &= ~flat
        done = active & (np.abs(p - p1) <= tol)
        root[done] = p[done]
        active &= ~done
        if not active.any():
            break
        active &= np.isfinite(p)            # a NaN iterate never recovers (scipy runs out of iterations and raises)
        if not active.any():
            break
        p0, q0 = p1, q1
        p1 = np.where(active, p, p1)
        q1 = f(p1)
    return root


def hydro_mass(pars, fit, r_kpc, cosmo, overdens=True, delta=500, start_opt=700.):
    """Hydrostatic mass profile and optionally ``(r_delta, m_delta)`` (``joxsz_plots.py:316-339``) for one
    parameter vector or a batch ``[S, ndim]``."""
    vals = np.asarray(pars, dtype=np.float64)
    single = vals.ndim == 1
    if single:
        fit.updateThawed(vals)
    bp = batched_pars(fit, np.atleast_2d(vals))
    m_prof = np.atleast_2d(fit.mass_cmpt.mass_fun(bp, r_kpc))
    if not overdens:
        return m_prof[0] if single else m_prof
    r_delta = overdensity_radius(np.atleast_2d(vals), fit, cosmo, delta, start_opt)
    m_delta = _mass_at(fit, bp, r_delta.reshape(-1, 1)).reshape(-1)
    if single:
        return m_prof[0], float(r_delta[0]), float(m_delta[0])
    return m_prof, r_delta, m_delta


def comp_mass_prof(cube_chain, fit, num="all", seed=None, overdens=True, delta=500, start_opt=700., ci=95):
    """Median and interval of the hydrostatic mass profile (and of ``r_delta``, ``m_delta``)
    (``joxsz_plots.py:341-376``)."""
    samples = select_samples(cube_chain, num, seed)
    res = hydro_mass(samples, fit, fit.data.sz.r_pp, fit.data.annuli.cosmology, overdens=overdens, delta=delta,
                     start_opt=start_opt)
    if overdens:
        m_prof, r_d, m_d = res
        return get_equal_tailed(m_prof, ci), get_equal_tailed(r_d, ci), get_equal_tailed(m_d, ci)
    return get_equal_tailed(res, ci)


def frac_gas_prof(cube_chain, fit, num="all", seed=None, ci=95):
    """Median and interval of the gas-fraction profile ``M_gas(<r) / M_HSE(<r)`` (``joxsz_plots.py:451-476``)."""
    samples = select_samples(cube_chain, num, seed)
    bp = batched_pars(fit, samples)
    r = fit.data.sz.r_pp
    dens = np.atleast_2d(fit.model.ne_cmpt.vikhFunction(bp, r))
    m_gas = cum_gas_mass(r, dens)
    m_tot = np.atleast_2d(fit.mass_cmpt.mass_fun(bp, r))
    return get_equal_tailed(m_gas / m_tot, ci)
