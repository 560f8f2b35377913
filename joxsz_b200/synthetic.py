"""Synthetic inputs: XSPEC-free count-rate tables, synthetic cluster geometries and parameter draws.

XSPEC/HEASOFT cannot run in this environment, so the per-band ``ln rate(ln T)`` tables that mbproj2
would build with ``phabs*apec`` (format: reference ``joxsz_funcs.py:652-681``) are replaced by smooth
bremsstrahlung-like curves.  They are *inputs*, shared verbatim by the CUDA path and the oracle; the
X-ray parity that follows from them is arithmetic parity (hard part 7).
"""
from __future__ import annotations

import numpy as np

# fiducial parameters inside every bound, in the thawed order of joxsz_main.py:179
FIDUCIAL = {
    "log(n_0)": -1.7, r"\beta": 0.67, "log(r_c)": 2.0, "log(r_s)": 2.7, r"\epsilon": 3.0,
    "log(T_X/T_{SZ})": 0.0, "Z": 0.3, "P_0": 0.25, "a": 1.33, "b": 4.13, "r_p": 300.0,
    "backscale": 1.0, "calibration": 1.0,
}


def synthetic_countrate_tables(bands_keV, Tlogvals, norm=3.0e-70):
    """``(ln rate_Z0, ln rate_Z1)`` per band on ``Tlogvals`` (natural log of keV).

    rate_Z0(T) = norm * T^-1/2 * (exp(-Emin/T) - exp(-Emax/T))   (thermal bremsstrahlung in the band)
    rate_Z1(T) = rate_Z0(T) * (1 + line bump centred near T ~ Emid/2)
    Floored at 1e-300 like the reference's cache builder (``joxsz_funcs.py:674``).
    """
    T = np.exp(np.asarray(Tlogvals, dtype=np.float64))
    out = []
    for emin, emax in bands_keV:
        r0 = norm * T ** -0.5 * (np.exp(-emin / T) - np.exp(-emax / T))
        emid = 0.5 * (emin + emax)
        bump = 1.0 + 2.5 * np.exp(-0.5 * ((np.log(T) - np.log(0.5 * emid)) / 0.9) ** 2)
        r1 = r0 * bump
        r0 = np.maximum(r0, 1e-300)
        r1 = np.maximum(r1, 1e-300)
        out.append((np.log(r0), np.log(r1)))
    return out


def draw_parameters(thawed, fiducial=None, n=1024, seed=20260102, spread=0.1, frac_bad=0.0, bounds=None):
    """``theta = fid * (1 + spread * N(0,1))`` per SURVEY.md section 8d, optionally with a fraction of
    walkers pushed out of bounds / to r_c > r_s to exercise the -inf paths."""
    fid = dict(FIDUCIAL if fiducial is None else fiducial)
    rng = np.random.default_rng(seed)
    base = np.array([fid[n_] for n_ in thawed], dtype=np.float64)
    theta = base[None, :] * (1.0 + spread * rng.standard_normal((n, base.size)))
    zero = base == 0.0
    if zero.any():   # multiplicative jitter cannot move a zero fiducial: use an additive one
        theta[:, zero] = spread * rng.standard_normal((n, int(zero.sum())))
    nbad = int(round(frac_bad * n))
    if nbad:
        idx = rng.choice(n, size=nbad, replace=False)
        for k, i in enumerate(idx):
            mode = k % 3
            if mode == 0 and "P_0" in thawed:                       # box prior violation
                theta[i, thawed.index("P_0")] = -0.05
            elif mode == 1 and "log(r_c)" in thawed and "log(r_s)" in thawed:   # r_c > r_s
                theta[i, thawed.index("log(r_c)")] = theta[i, thawed.index("log(r_s)")] + 0.3
            elif "b" in thawed:                                      # outside the box on another axis
                theta[i, thawed.index("b")] = 15.5
    return theta
