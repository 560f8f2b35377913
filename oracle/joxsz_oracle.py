"""CPU oracle for the JoXSZ per-walker joint SZ + X-ray log-likelihood.

TEST INFRASTRUCTURE ONLY.  This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it.  Nothing under ``joxsz_b200/`` imports it and it imports nothing from ``joxsz_b200/``.

It restates, function by function and in float64 numpy/scipy like the reference, the hot path of
fcastagna/JoXSZ (``/root/reference/joxsz_funcs.py``; line numbers below refer to that file unless
prefixed ``main:`` for ``joxsz_main.py``).

PARITY PINNING -- read this before trusting a number:

* PINNED against the reference's own code: ``tests/golden/make_golden_reference.py`` imports the
  unmodified ``/root/reference/joxsz_funcs.py`` (third-party imports stubbed) and records
  ``getLikelihood`` / ``get_sz_like('pp'|'bright'|'chisq'|'ll')`` / ``calcProfiles`` outputs on seeded
  parameter draws; ``tests/test_oracle_golden.py`` checks this oracle against those vectors.  This
  pins everything that lives in the reference: gNFW pressure and derivative, Vikhlinin density,
  temperatures, HSE-mass veto, the SZ staging with scipy's interp1d/fftconvolve/fft2, the Cash sum,
  the -inf ordering in ``getLikelihood``.
* UNPINNED ("parity unpinned") at the third-party boundary: PyAbel's ``direct_transform``, mbproj2's
  ``Param*.prior`` / ``Annuli`` / ``Band.calcProjProfile`` / ``CountRate.getCountRate`` /
  ``cashLogLikelihood`` / ``Cosmology`` and emcee's stretch move are NOT under ``/root/reference``,
  are unpinned in ``requirements.txt`` and cannot be installed here; XSPEC (which builds mbproj2's
  count-rate tables) is absent.  They are restated from their published algorithms (SURVEY.md
  Appendix A) and anchored by closed-form known-answer tests (Abel pairs, projection-volume sums,
  hand-computed Cash/priors) in ``tests/test_oracle_kat.py``.  The golden generator necessarily uses
  these same restatements as its stubs, so the golden vectors do not pin them.

Two formulations are provided and cross-checked: the literal staged path (per walker, operators
rebuilt per call as PyAbel/scipy do -- this is also the "reference CPU path" timed by the bench) and a
batched path with the fixed linear operators precomputed.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.fftpack import fft2, ifft2
from scipy.interpolate import interp1d
from scipy.signal import fftconvolve

# mbproj2.physconstants [MEM]; only kpc_cm changes hot-path numbers.
KPC_CM = 3.0856776e21
KEV_ERG = 1.6022e-9
MU_G = 1.6605e-24
G_CGS = 6.67428e-8
SOLAR_MASS_G = 1.989e33

# canonical names of the JoXSZ parameters (defPars: 256-273, 318, 358-373; main:131,156-157)
N_P0, N_A, N_B, N_C, N_RP = "P_0", "a", "b", "c", "r_p"
N_LN0, N_BETA, N_LRC, N_LRS = "log(n_0)", r"\beta", "log(r_c)", "log(r_s)"
N_ALPHA, N_EPS, N_GAMMA = r"\alpha", r"\epsilon", r"\gamma"
N_LN02, N_BETA2, N_LRC2 = "log(n_{02})", r"\beta_2", "log(r_{c2})"
N_LTR, N_Z, N_BACK, N_CAL = "log(T_X/T_{SZ})", "Z", "backscale", "calibration"


# ----------------------------------------------------------------------------------------
# physics components (dict p: name -> float)
# ----------------------------------------------------------------------------------------

def press_fun(p, r_kpc):
    """gNFW pressure, ``CmptPressure.press_fun`` (275-287)."""
    P_0, r_p, a, b, c = p[N_P0], p[N_RP], p[N_A], p[N_B], p[N_C]
    return P_0 / ((r_kpc / r_p) ** c * (1 + (r_kpc / r_p) ** a) ** ((b - c) / a))


def press_derivative(p, r_kpc):
    """dP/dr, ``CmptPressure.press_derivative`` (289-301)."""
    P_0, r_p, a, b, c = p[N_P0], p[N_RP], p[N_A], p[N_B], p[N_C]
    return -P_0 * (c + b * (r_kpc / r_p) ** a) / (
        r_p * (r_kpc / r_p) ** (c + 1) * (1 + (r_kpc / r_p) ** a) ** ((b - c + a) / a))


def vikh_density(p, radii_kpc, mode="single"):
    """Vikhlinin density, ``mydens_vikhFunction`` (375-395)."""
    n_0 = 10 ** p[N_LN0]
    beta = p[N_BETA]
    r_c = 10 ** p[N_LRC]
    r_s = 10 ** p[N_LRS]
    alpha, epsilon, gamma = p[N_ALPHA], p[N_EPS], p[N_GAMMA]
    r = radii_kpc
    res_sq = n_0 ** 2 * (r / r_c) ** (-alpha) / (
        (1 + (r / r_c) ** 2) ** (3 * beta - alpha / 2) * (1 + (r / r_s) ** gamma) ** (epsilon / gamma))
    if mode == "double":
        n_02 = 10 ** p[N_LN02]
        r_c2 = 10 ** p[N_LRC2]
        beta_2 = p[N_BETA2]
        res_sq = res_sq + n_02 ** 2 / (1 + (r / r_c2) ** 2) ** (3 * beta_2)
    return np.sqrt(res_sq)


def dens_prior(p):
    """``mydens_prior`` (397-407): -inf when r_c > r_s."""
    if 10 ** p[N_LRC] > 10 ** p[N_LRS]:
        return -np.inf
    return 0.0


def temp_fun(p, r_kpc, mode="single", getT_SZ=False):
    """``CmptUPPTemperature.temp_fun`` (321-336)."""
    T_SZ = press_fun(p, r_kpc) / vikh_density(p, r_kpc, mode)
    if getT_SZ:
        return T_SZ
    return T_SZ * 10 ** p[N_LTR]


def mass_fun(p, r_kpc, mode="single", mu_gas=0.61):
    """Hydrostatic mass, ``CmptMyMass.mass_fun`` (428-437)."""
    dpr_cm = press_derivative(p, r_kpc) * KEV_ERG / KPC_CM
    ne = vikh_density(p, r_kpc, mode)
    r_cm = r_kpc * KPC_CM
    return -dpr_cm * r_cm ** 2 / (mu_gas * MU_G * ne * G_CGS) / SOLAR_MASS_G


def mass_is_monotone(p, r_kpc, mode="single"):
    """Veto of 523-525: ``all(np.gradient(m_prof, 1) > 0.)``."""
    with np.errstate(all="ignore"):
        return bool(np.all(np.gradient(mass_fun(p, r_kpc, mode), 1) > 0.0))


# ----------------------------------------------------------------------------------------
# PyAbel direct transform, Python backend  [third party, restated: SURVEY.md Appendix A.1]
# ----------------------------------------------------------------------------------------

_trapz = getattr(np, "trapezoid", None) or np.trapz


def _is_uniform_sampling(r):
    return bool(np.allclose(np.diff(np.diff(r)), 0.0, atol=1e-13))


def _pyabel_direct_integral(f, r, correction):
    """``abel.direct._pyabel_direct_integral``: trapezoid rule off the singular cell, minus half of
    the spike trapezoids at j=i+1, plus the analytic first cell for a piecewise-linear integrand."""
    if _is_uniform_sampling(r):
        int_opts = {"dx": abs(r[1] - r[0])}
    else:
        int_opts = {"x": r}
    out = np.zeros(f.shape)
    R, Y = np.meshgrid(r, r, indexing="ij")
    i_vect = np.arange(len(r), dtype=int)
    II, JJ = np.meshgrid(i_vect, i_vect, indexing="ij")
    mask = II < JJ
    I_sqrt = np.zeros(R.shape)
    I_sqrt[mask] = np.sqrt((Y ** 2 - R ** 2)[mask])
    I_isqrt = np.zeros(R.shape)
    I_isqrt[mask] = 1.0 / I_sqrt[mask]
    mask2 = (II > JJ - 2) & (II < JJ + 1)
    for i, row in enumerate(f):
        P = row[None, :] * I_isqrt
        out[i, :] = _trapz(P, axis=1, **int_opts)
        out[i, :] = out[i, :] - 0.5 * _trapz(P * mask2, axis=1, **int_opts)
    if correction == 1:
        f_r = (f[:, 1:] - f[:, :-1]) / np.diff(r)[None, :]
        isqrt = I_sqrt[II + 1 == JJ]
        if r[0] < r[1] * 1e-8:
            ratio = np.append(np.cosh(1), r[2:] / r[1:-1])
        else:
            ratio = r[1:] / r[:-1]
        acr = np.arccosh(ratio)
        for i, row in enumerate(f):
            out[i, :-1] += isqrt * f_r[i] + acr * (row[:-1] - f_r[i] * r[:-1])
    return out


def pyabel_direct_forward(fr, r):
    """``abel.direct.direct_transform(fr, r=r, direction='forward', backend='Python')`` (call: 457)."""
    f = np.atleast_2d(np.array(fr, dtype=np.float64, copy=True))
    f = f * (2 * r[None, :])
    out = _pyabel_direct_integral(f, np.asarray(r, dtype=np.float64), 1)
    return out[0] if np.ndim(fr) == 1 else out


# ----------------------------------------------------------------------------------------
# set-up container (plain arrays only)
# ----------------------------------------------------------------------------------------

class OracleSetup:
    """Constants of one cluster set-up as plain numpy arrays / floats.

    SZ (names follow ``SZ_data``, 136-170): ``phys_const`` [m_e keV, sigma_T cm2], ``step``, ``kpc_as``,
    ``conv_T``/``conv_I`` (the y->mJy/beam table, main:108-109, I already x1e3), ``flux_data`` [3,Nd],
    ``beam_2d``, ``radius``, ``sep``, ``r_pp``, ``d_mat``, ``filtering``, ``calc_integ``, ``integ_mu``, ``integ_sig``.
    X: ``midpt_kpc`` [Na], ``projvols_cm3`` [Na,Na] (annulus x shell), ``geomarea_arcmin2`` [Na],
    ``bands`` = list of dict(cts, areascales, exposures, backrates, lnrate_Z0, lnrate_Z1), ``Tlogvals``,
    ``Tmin``, ``Tmax``.
    Parameters: ``par_names`` (all, dict order), ``par_kind`` ('box'|'gauss'), ``par_a``/``par_b``
    (min,max | mu,sigma), ``par_val`` (current/frozen values), ``thawed`` (names, sampling order),
    ``dens_mode``, ``exclude_unphy_mass``.
    """

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def convert(self, T):
        """``interp1d(t_keV, 1e3*I0, 'linear', fill_value='extrapolate')`` (main:109)."""
        return interp1d(self.conv_T, self.conv_I, "linear", fill_value="extrapolate")(T)

    def full_params(self, theta):
        """``Fit.updateThawed`` (516): scatter thawed values over the frozen ones."""
        p = dict(zip(self.par_names, (float(v) for v in self.par_val)))
        for name, v in zip(self.thawed, theta):
            p[name] = float(v)
        return p


# ----------------------------------------------------------------------------------------
# literal staged path, one walker
# ----------------------------------------------------------------------------------------

def param_prior(p, s):
    """``sum(self.pars[p].prior() for p in self.pars)`` (518): box -> 0/-inf, Gaussian -> log-pdf."""
    total = 0.0
    for name, kind, a, b in zip(s.par_names, s.par_kind, s.par_a, s.par_b):
        v = p[name]
        if kind == "box":
            if v < a or v > b:
                total += -np.inf
        else:
            if b > 0:
                total += -0.5 * math.log(2 * math.pi) - math.log(b) - 0.5 * ((v - a) / b) ** 2
    return total


def sz_stages(p, s):
    """``get_sz_like`` (439-493) with every intermediate kept."""
    out = {}
    pp = press_fun(p, s.r_pp)                                                    # 453
    out["pp"] = pp
    ab = pyabel_direct_forward(pp, s.r_pp)                                       # 457
    out["ab"] = ab
    y = KPC_CM * s.phys_const[1] / s.phys_const[0] * ab                          # 459
    out["y"] = y
    f = interp1d(np.append(-s.r_pp, s.r_pp), np.append(y, y), "cubic",
                 bounds_error=False, fill_value=(0.0, 0.0))                      # 460
    y_2d = f(s.d_mat)                                                            # 462
    out["y_2d"] = y_2d
    conv_2d = fftconvolve(y_2d, s.beam_2d, "same") * s.step ** 2                 # 464
    out["conv_2d"] = conv_2d
    map_out = np.real(ifft2(fft2(conv_2d) * s.filtering))                        # 466-467
    out["map_out"] = map_out
    r_t = s.r_pp[:s.sep]
    t_prof = temp_fun(p, r_t, s.dens_mode, getT_SZ=True)                         # 469
    out["t_prof"] = t_prof
    h = interp1d(np.append(-r_t, r_t), np.append(t_prof, t_prof), "cubic",
                 bounds_error=False, fill_value=(t_prof[-1], t_prof[-1]))        # 470-471
    c = conv_2d.shape[0] // 2
    t_all = np.append(h(0.0), t_prof)
    out["T0"] = float(t_all[0])
    map_prof = map_out[c, c:] * s.convert(t_all) * p[N_CAL]                      # 472-473
    out["bright"] = map_prof
    g = interp1d(s.radius[s.sep:], map_prof, "cubic", fill_value="extrapolate")  # 476
    model = g(s.flux_data[0])
    out["model"] = model
    chisq = np.nansum(((s.flux_data[1] - model) / s.flux_data[2]) ** 2)          # 478
    log_lik = -chisq / 2                                                         # 479
    if s.calc_integ:                                                             # 480-485
        from scipy.integrate import simpson
        x = np.arange(0.0, s.r_pp[-1] / s.kpc_as / 60 + s.step / 60, s.step / 60)
        cint = simpson(np.concatenate((f(0.0), y), axis=None) * x, x=x) * 2 * np.pi
        out["integ"] = cint
        log_lik -= np.nansum(((cint - s.integ_mu) / s.integ_sig) ** 2) / 2
    out["chisq"] = chisq
    out["ll"] = log_lik
    return out


def xray_profiles(p, s):
    """``Fit.calcProfiles`` -> ``ModelNullPot.computeProfs`` -> ``Band.calcProjProfile`` with
    ``CountRate.getCountRate``  [mbproj2, restated: SURVEY.md Appendix A.3]; call site 527."""
    ne = vikh_density(p, s.midpt_kpc, s.dens_mode)
    T = temp_fun(p, s.midpt_kpc, s.dens_mode)                                    # 338-339
    Z = np.full(s.midpt_kpc.size, p[N_Z])
    with np.errstate(all="ignore"):
        logT = np.log(np.clip(T, s.Tmin, s.Tmax))
    profs = []
    for band in s.bands:
        r0 = np.exp(np.interp(logT, s.Tlogvals, band["lnrate_Z0"]))
        r1 = np.exp(np.interp(logT, s.Tlogvals, band["lnrate_Z1"]))
        rates = (r0 + (r1 - r0) * Z) * ne ** 2
        proj = s.projvols_cm3.dot(rates) * (band["areascales"] * band["exposures"])
        proj = proj + (band["backrates"] * p[N_BACK] * s.geomarea_arcmin2
                       * band["areascales"] * band["exposures"])
        profs.append(proj)
    return profs


def cash_log_likelihood(data, model):
    """``mbproj2.utils.cashLogLikelihood`` [restated]: sum(d ln m) - sum(m); -inf if not finite."""
    with np.errstate(all="ignore"):
        like = np.sum(data * np.log(model)) - np.sum(model)
    return like if np.isfinite(like) else -np.inf


def xray_like_from_profs(profs, s):
    """``mylikeFromProfs`` (495-505): Cash over bins whose counts are not NaN."""
    likelihood = 0.0
    for band, pred in zip(s.bands, profs):
        ok = ~np.isnan(band["cts"])
        likelihood += cash_log_likelihood(band["cts"][ok], pred[ok])
    return likelihood


def get_likelihood(theta, s, detail=False):
    """``getLikelihood`` (507-546) for one walker; same early returns, same order."""
    p = s.full_params(theta)                                                     # 515-516
    parprior = param_prior(p, s)                                                 # 518
    if not np.isfinite(parprior):                                                # 519-520
        return (-np.inf, {"why": "prior"}) if detail else -np.inf
    if s.exclude_unphy_mass:                                                     # 522-525
        if not mass_is_monotone(p, s.r_pp, s.dens_mode):
            return (-np.inf, {"why": "mass"}) if detail else -np.inf
    profs = xray_profiles(p, s)                                                  # 527
    if np.array(profs).min() > 0.0:                                              # 529-532
        like = xray_like_from_profs(profs, s)
    else:
        like = -np.inf
    with np.errstate(all="ignore"):
        st = sz_stages(p, s)                                                     # 534
    prior = dens_prior(p) + parprior                                             # 536 (T and Z cmpt priors are 0)
    totlike = float(like + prior + st["ll"])                                     # 538
    if detail:
        st.update(xprofs=np.array(profs), xlike=like, prior=prior, why="ok")
        return totlike, st
    return totlike


def get_likelihood_many(thetas, s):
    """Literal path over a batch (a plain loop, one task per walker as emcee's pool.map does)."""
    return np.array([get_likelihood(t, s) for t in np.atleast_2d(thetas)])


# ----------------------------------------------------------------------------------------
# batched path with precomputed linear operators
# ----------------------------------------------------------------------------------------

class BatchedOracle:
    """Same arithmetic with the geometry-only work hoisted out of the walker loop.

    ``A`` = Abel matrix obtained by pushing unit vectors through :func:`pyabel_direct_forward`;
    ``L`` = the whole linear SZ chain pressure -> filtered map row (``map_out[c, c:]``), obtained by
    pushing unit vectors through the literal scipy stages (the "collapsed operator" cross-check of
    SURVEY.md section 4).  Used to pin the staged path against itself and as the fast CPU baseline.
    """

    def __init__(self, s: OracleSetup):
        self.s = s
        nr = s.r_pp.size
        self.A = pyabel_direct_forward(np.eye(nr), s.r_pp).T            # ab = A @ pp
        self.yscale = KPC_CM * s.phys_const[1] / s.phys_const[0]
        c = s.d_mat.shape[0] // 2
        rows = np.empty((nr, s.d_mat.shape[0] - c))
        for k in range(nr):
            e = np.zeros(nr)
            e[k] = 1.0
            f = interp1d(np.append(-s.r_pp, s.r_pp), np.append(e, e), "cubic",
                         bounds_error=False, fill_value=(0.0, 0.0))
            conv = fftconvolve(f(s.d_mat), s.beam_2d, "same") * s.step ** 2
            rows[k] = np.real(ifft2(fft2(conv) * s.filtering))[c, c:]
        self.M = rows.T                                                  # row = M @ y
        self.L = self.M @ (self.yscale * self.A)                         # row = L @ pp

    def _cols(self, thetas):
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        s = self.s
        cols = {n: np.full(thetas.shape[0], float(v)) for n, v in zip(s.par_names, s.par_val)}
        for j, n in enumerate(s.thawed):
            cols[n] = thetas[:, j].copy()
        return {n: v[:, None] for n, v in cols.items()}

    def map_row(self, thetas):
        p = self._cols(thetas)
        return press_fun(p, self.s.r_pp[None, :]) @ self.L.T

    def loglike(self, thetas):
        """Batched joint log-likelihood; agrees with :func:`get_likelihood` walker by walker."""
        s = self.s
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        W = thetas.shape[0]
        p = self._cols(thetas)
        with np.errstate(all="ignore"):
            parprior = np.zeros(W)
            for name, kind, a, b in zip(s.par_names, s.par_kind, s.par_a, s.par_b):
                v = p[name][:, 0]
                if kind == "box":
                    parprior = np.where((v < a) | (v > b), -np.inf, parprior)
                elif b > 0:
                    parprior = parprior - 0.5 * math.log(2 * math.pi) - math.log(b) - 0.5 * ((v - a) / b) ** 2
            dead = ~np.isfinite(parprior)
            r = s.r_pp[None, :]
            if s.exclude_unphy_mass:
                m = mass_fun(p, r, s.dens_mode)
                dead |= ~np.all(np.gradient(m, 1, axis=1) > 0.0, axis=1)
            # X-ray
            ra = s.midpt_kpc[None, :]
            ne = vikh_density(p, ra, s.dens_mode)
            T = temp_fun(p, ra, s.dens_mode)
            logT = np.log(np.clip(T, s.Tmin, s.Tmax))
            like = np.zeros(W)
            minprof = np.full(W, np.inf)
            for band in s.bands:
                r0 = np.exp(np.interp(logT, s.Tlogvals, band["lnrate_Z0"]))
                r1 = np.exp(np.interp(logT, s.Tlogvals, band["lnrate_Z1"]))
                rates = (r0 + (r1 - r0) * p[N_Z]) * ne ** 2
                sc = band["areascales"] * band["exposures"]
                pred = rates @ s.projvols_cm3.T * sc + band["backrates"] * p[N_BACK] * s.geomarea_arcmin2 * sc
                minprof = np.fmin(minprof, np.where(np.isnan(pred), -np.inf, pred).min(axis=1))
                ok = ~np.isnan(band["cts"])
                lb = (band["cts"][ok] * np.log(pred[:, ok])).sum(axis=1) - pred[:, ok].sum(axis=1)
                like = like + np.where(np.isfinite(lb), lb, -np.inf)
            like = np.where(minprof > 0.0, like, -np.inf)
            # SZ
            row = press_fun(p, r) @ self.L.T
            r_t = s.r_pp[:s.sep]
            t_prof = temp_fun(p, r_t[None, :], s.dens_mode, getT_SZ=True)
            sz = np.empty(W)
            for w in range(W):
                h = interp1d(np.append(-r_t, r_t), np.append(t_prof[w], t_prof[w]), "cubic")
                prof = row[w] * s.convert(np.append(h(0.0), t_prof[w])) * p[N_CAL][w, 0]
                g = interp1d(s.radius[s.sep:], prof, "cubic", fill_value="extrapolate")
                sz[w] = -np.nansum(((s.flux_data[1] - g(s.flux_data[0])) / s.flux_data[2]) ** 2) / 2
            dprior = np.where(10 ** p[N_LRC][:, 0] > 10 ** p[N_LRS][:, 0], -np.inf, 0.0)
            tot = like + (dprior + parprior) + sz
        return np.where(dead, -np.inf, tot)
