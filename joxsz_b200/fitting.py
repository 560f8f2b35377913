"""Batched replacement for ``Fit.doFitting`` (called at reference ``joxsz_main.py:191``).

mbproj2's ``doFitting`` alternates scipy's serial Nelder-Mead / Powell minimisers on ``-getLikelihood``
until the likelihood improves by less than 0.1: thousands of likelihood evaluations, one at a time.
Here the same criterion drives a *batched multi-start Nelder-Mead*: ``nstarts`` independent simplices
advance together, and every iteration evaluates the reflection, expansion and both contraction points of
all simplices in ONE ``jx_loglike`` call (``4 * nstarts`` walkers); shrinks, which are rare, take a second
call.  Coefficients, initial simplex (5 % / 0.00025 perturbations) and the ``xatol = fatol = 1e-4`` stopping
rule are scipy's ``minimize(method='Nelder-Mead')`` defaults, so start 0 follows the trajectory the serial
minimiser would take from the current parameters; the other starts are jittered copies that guard against
the local maxima the -inf regions of this likelihood create.

Host side = simplex bookkeeping on ``[nstarts, ndim + 1, ndim]`` numpy arrays; all likelihood arithmetic is
on the GPU.
"""
from __future__ import annotations

import numpy as np

from .funcs import engine_for

_BIG = 1e99          # value given to -inf / NaN likelihoods, like mbproj2's minimiser wrapper


def _neg_loglike(eng, x):
    """f = -loglike for points x [..., ndim] (any leading shape), evaluated in engine-sized batches."""
    flat = np.ascontiguousarray(x.reshape(-1, x.shape[-1]))
    out = np.empty(flat.shape[0])
    step = eng.max_walkers
    for lo in range(0, flat.shape[0], step):
        out[lo:lo + step] = eng(flat[lo:lo + step])
    f = -out
    f[~np.isfinite(f)] = _BIG
    return f.reshape(x.shape[:-1])


def initial_simplices(x0, nstarts, spread, rng):
    """[nstarts, ndim+1, ndim]: scipy's default simplex around each start; start 0 is exactly x0."""
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.size
    starts = np.tile(x0, (nstarts, 1))
    if nstarts > 1:
        jit = x0 * spread * rng.standard_normal((nstarts - 1, n))
        jit[:, x0 == 0.0] = spread * rng.standard_normal((nstarts - 1, int((x0 == 0.0).sum())))
        starts[1:] += jit
    sim = np.repeat(starts[:, None, :], n + 1, axis=1)
    for k in range(n):
        col = sim[:, k + 1, k]
        sim[:, k + 1, k] = np.where(col != 0.0, 1.05 * col, 0.00025)
    return sim


def batched_nelder_mead(eng, sim, maxiter=None, xatol=1e-4, fatol=1e-4):
    """Minimise -loglike from the simplices ``sim`` [M, n+1, n].  Returns (x_best [M, n], f_best [M], nfev)."""
    M, n1, n = sim.shape
    rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    if maxiter is None:
        maxiter = 200 * n
    sim = sim.copy()
    f = _neg_loglike(eng, sim)
    nfev = f.size
    rows = np.arange(M)
    active = np.ones(M, dtype=bool)
    for _ in range(maxiter):
        order = np.argsort(f, axis=1, kind="stable")
        f = np.take_along_axis(f, order, axis=1)
        sim = np.take_along_axis(sim, order[:, :, None], axis=1)
        conv = (np.max(np.abs(sim[:, 1:] - sim[:, :1]), axis=(1, 2)) <= xatol) & \
               (np.max(np.abs(f[:, 1:] - f[:, :1]), axis=1) <= fatol)
        active &= ~conv
        if not active.any():
            break
        xbar = sim[:, :-1].mean(axis=1)
        xw = sim[:, -1]
        cand = np.stack([(1 + rho) * xbar - rho * xw,                    # reflection
                         (1 + rho * chi) * xbar - rho * chi * xw,        # expansion
                         (1 + psi * rho) * xbar - psi * rho * xw,        # outside contraction
                         (1 - psi) * xbar + psi * xw], axis=1)           # inside contraction
        fc = np.full((M, 4), _BIG)
        fc[active] = _neg_loglike(eng, cand[active])
        nfev += 4 * int(active.sum())
        fr, fe, foc, fic = fc.T
        fbest, fsecond, fworst = f[:, 0], f[:, -2], f[:, -1]
        choice = np.full(M, -1)                                           # index into cand; -1 = shrink
        c1 = fr < fbest
        choice[c1 & (fe < fr)] = 1
        choice[c1 & ~(fe < fr)] = 0
        c2 = ~c1 & (fr < fsecond)
        choice[c2] = 0
        c3 = ~c1 & ~c2 & (fr < fworst)
        choice[c3 & (foc <= fr)] = 2
        c4 = ~c1 & ~c2 & ~c3
        choice[c4 & (fic < fworst)] = 3
        take = active & (choice >= 0)
        sim[take, -1] = cand[take, choice[take]]
        f[take, -1] = fc[take, choice[take]]
        shrink = active & (choice < 0)
        if shrink.any():
            sim[shrink, 1:] = sim[shrink, :1] + sigma * (sim[shrink, 1:] - sim[shrink, :1])
            f[shrink, 1:] = _neg_loglike(eng, sim[shrink, 1:])
            nfev += int(shrink.sum()) * n
    best = np.argmin(f, axis=1)
    return sim[rows, best], f[rows, best], nfev


def doFitting(self, silent=False, maxiter=10, nstarts=64, spread=0.02, seed=0):
    """``Fit.doFitting`` work-alike: maximise the joint likelihood from the current thawed parameters,
    repeating multi-start simplex rounds until the best likelihood improves by less than 0.1 (mbproj2's
    stopping rule), then leave the fit at the best point.  Returns the best log-likelihood."""
    eng = engine_for(self, 4 * nstarts)
    rng = np.random.default_rng(seed)
    x = np.array(self.thawedParVals(), dtype=np.float64)
    like = float(eng(x))
    for it in range(maxiter):
        sim = initial_simplices(x, nstarts, spread, rng)
        xb, fb, nfev = batched_nelder_mead(eng, sim)
        k = int(np.argmin(fb))
        newlike = -float(fb[k])
        if newlike > like or not np.isfinite(like):
            x = xb[k]
        else:
            newlike = like
        if not silent:
            print("Fit: %g -> %g  (%d likelihood evaluations in %d-walker batches)" % (like, newlike, nfev, 4 * nstarts))
        done = np.isfinite(like) and abs(newlike - like) < 0.1
        like = newlike
        if done:
            break
    self.updateThawed(x)
    self.bestlike = max(getattr(self, "bestlike", -1e99), like)
    return like
