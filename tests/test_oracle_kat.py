"""Closed-form known-answer tests for the third-party pieces the oracle restates from memory
(PyAbel direct transform, mbproj2 projection volumes / Cash / priors) and for the scipy semantics the
reference relies on."""
import math

import numpy as np
from scipy.interpolate import interp1d, CubicSpline
from scipy.signal import fftconvolve, convolve2d
from scipy.special import gamma as Gamma

from helpers import orc


def test_abel_gaussian_pair():
    # f = exp(-r^2/s^2)  ->  F(y) = s sqrt(pi) exp(-y^2/s^2)
    h, s = 16.0, 400.0
    r = h * np.arange(1, 314)
    F = orc.pyabel_direct_forward(np.exp(-(r / s) ** 2), r)
    exact = s * math.sqrt(math.pi) * np.exp(-(r / s) ** 2)
    inner = r < 600.0
    assert np.max(np.abs(F[inner] / exact[inner] - 1)) < 1e-2      # discretisation error of the method itself
    assert abs(F[0] / exact[0] - 1) < 1e-3
    assert F[-1] == 0.0


def test_abel_beta_model_pair():
    # (1 + r^2/rc^2)^(-3b/2) -> sqrt(pi) G(3b/2-1/2)/G(3b/2) rc (1 + y^2/rc^2)^(-3b/2+1/2)
    h, rc, b = 4.0, 120.0, 1.2
    r = h * np.arange(1, 2501)
    F = orc.pyabel_direct_forward((1 + (r / rc) ** 2) ** (-1.5 * b), r)
    exact = math.sqrt(math.pi) * Gamma(1.5 * b - 0.5) / Gamma(1.5 * b) * rc * (1 + (r / rc) ** 2) ** (-1.5 * b + 0.5)
    inner = r < 1000.0
    assert np.max(np.abs(F[inner] / exact[inner] - 1)) < 1e-2


def test_abel_is_linear_and_batch_consistent():
    rng = np.random.default_rng(1)
    r = 16.00139 * np.arange(1, 200)
    a, b = rng.random(r.size), rng.random(r.size)
    Fa, Fb = orc.pyabel_direct_forward(a, r), orc.pyabel_direct_forward(b, r)
    Fab = orc.pyabel_direct_forward(2 * a - 3 * b, r)
    assert np.max(np.abs(Fab - (2 * Fa - 3 * Fb))) < 1e-9 * np.max(np.abs(Fab))
    both = orc.pyabel_direct_forward(np.stack([a, b]), r)
    np.testing.assert_allclose(both[0], Fa, rtol=1e-13)


def test_cubic_interp1d_is_notaknot_and_reproduces_cubics():
    x = np.sort(np.random.default_rng(2).uniform(-3, 3, 30))
    y = 1 + 2 * x - 0.5 * x ** 2 + 0.25 * x ** 3
    xq = np.linspace(-3.5, 3.5, 101)
    f = interp1d(x, y, "cubic", fill_value="extrapolate")
    np.testing.assert_allclose(f(xq), 1 + 2 * xq - 0.5 * xq ** 2 + 0.25 * xq ** 3, rtol=1e-10, atol=1e-10)
    yr = np.cos(x)
    np.testing.assert_allclose(interp1d(x, yr, "cubic")(x[3:-3] + 0.01),
                               CubicSpline(x, yr, bc_type="not-a-knot")(x[3:-3] + 0.01), rtol=1e-12, atol=1e-13)
    assert np.all(np.isnan(interp1d(x, np.where(np.arange(30) == 7, np.nan, yr), "cubic")(xq[20:80])))


def test_fftconvolve_same_is_centred_linear_convolution():
    rng = np.random.default_rng(3)
    a, k = rng.random((21, 21)), rng.random((7, 7))
    np.testing.assert_allclose(fftconvolve(a, k, "same"), convolve2d(a, k, "same"), atol=1e-13)


def test_projection_volumes_sum_to_shell_volumes():
    import sys
    from joxsz_b200.mbshim import utils
    edges = np.array([0, .05, .1, .15, .2, .25, .3, .4, .5, 1, 1.3333, 2, 2.6667, 4.3333, 6, 7.6667]) * 480.0
    m = utils.projectionVolumeMatrix(edges)              # [shell, annulus]
    np.testing.assert_allclose(m.sum(axis=1), 4.0 / 3.0 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3), rtol=1e-12)
    assert np.all(np.triu(m, k=1) == 0.0)                # a shell only projects onto annuli inside it
    assert np.all(m >= 0.0)


def test_cash_and_priors_hand_values(cl1226_oracle):
    d, m = np.array([3.0, 0.0, 5.0]), np.array([2.0, 1.5, 4.0])
    assert abs(orc.cash_log_likelihood(d, m) - (3 * math.log(2) + 5 * math.log(4) - 7.5)) < 1e-14
    assert orc.cash_log_likelihood(d, np.array([2.0, -1.0, 4.0])) == -np.inf
    s = cl1226_oracle
    p = s.full_params(np.array([dict(zip(s.par_names, s.par_val))[n] for n in s.thawed]))
    base = orc.param_prior(p, s)
    # two Gaussian priors at their means: -ln(sigma sqrt(2 pi)) each
    expect = -(math.log(0.1) + math.log(0.07)) - math.log(2 * math.pi)
    assert abs(base - expect) < 1e-13
    p2 = dict(p); p2["P_0"] = 2.5
    assert orc.param_prior(p2, s) == -np.inf
    p3 = dict(p); p3["log(r_c)"] = 3.0; p3["log(r_s)"] = 2.0
    assert orc.dens_prior(p3) == -np.inf and orc.dens_prior(p) == 0.0


def test_filter_identity_and_nan_quirk(cl1226_oracle):
    s = cl1226_oracle
    import copy
    s1 = copy.copy(s)
    s1.filtering = np.ones_like(s.filtering)
    p = s.full_params([dict(zip(s.par_names, s.par_val))[n] for n in s.thawed])
    st = orc.sz_stages(p, s1)
    assert np.max(np.abs(st["map_out"] - st["conv_2d"])) < 1e-12 * np.max(np.abs(st["conv_2d"]))
    # NaN anywhere in the profile -> every residual NaN -> nansum gives chi^2 = 0 (finite ll)
    p["P_0"] = float("nan")
    with np.errstate(all="ignore"):
        assert orc.sz_stages(p, s)["chisq"] == 0.0


def test_abel_operator_is_pinned_to_its_discretisation():
    """Tight KAT for the PyAbel direct transform as the reference calls it (joxsz_funcs.py:457): an independent scalar
    restatement of the discretisation -- first cell integrated exactly for the piecewise-linear integrand (here by
    adaptive quadrature of the substituted, smooth integrand), trapezoid rule on f(r)/sqrt(r^2 - y^2) beyond it, last
    point zero -- must agree with the oracle and with the dense operator the CUDA path uses to 1e-12.  The closed-form
    pairs above only bound the method's own discretisation error (1e-2 .. 1e-3): a different but convergent variant
    would pass them and fail this one."""
    from scipy.integrate import quad
    from joxsz_b200 import operators as ops
    rng = np.random.default_rng(5)
    r = 16.00139 * np.arange(1, 61)
    pp = np.exp(-r / 300.0) * (1 + 0.1 * rng.standard_normal(r.size))
    f = 2.0 * r * pp
    n = r.size
    want = np.zeros(n)
    for i in range(n - 1):
        y = r[i]
        slope = (f[i + 1] - f[i]) / (r[i + 1] - r[i])
        # integral over [r_i, r_{i+1}] of (f_i + slope (r - r_i)) / sqrt(r^2 - y^2) dr with r = y cosh(t)
        first, _ = quad(lambda t: f[i] + slope * (y * math.cosh(t) - y), 0.0, math.acosh(r[i + 1] / y),
                        epsabs=0, epsrel=2e-14)
        g = f[i + 1:] / np.sqrt(r[i + 1:] ** 2 - y * y)
        trap = float(np.sum(0.5 * (g[1:] + g[:-1]) * np.diff(r[i + 1:])))
        if i == n - 2:
            # edge artefact of PyAbel's code, kept because the reference calls that code: the half-weight it takes off
            # the spike sample j = i + 1 is computed from the two intervals around it, and the last sample has only
            # one -- a quarter of the spike trapezoid survives in the second-to-last output
            trap += 0.25 * g[0] * (r[i + 1] - r[i])
        want[i] = first + trap
    got = orc.pyabel_direct_forward(pp, r)
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) < 1e-12 * scale
    assert got[-1] == 0.0
    A = ops.abel_forward_matrix(r)
    assert np.max(np.abs(A @ pp - want)) < 1e-12 * scale
    # convergence order of this discretisation against the closed form: the trapezoid rule meets the inverse-square-root
    # behaviour next to the exactly integrated first cell, so halving the step divides the error by ~sqrt(2), not 4
    s = 400.0
    errs = []
    for h in (16.0, 8.0, 4.0):
        rr = h * np.arange(1, int(5000 / h) + 1)
        F = orc.pyabel_direct_forward(np.exp(-(rr / s) ** 2), rr)
        exact = s * math.sqrt(math.pi) * np.exp(-(rr / s) ** 2)
        k = int(round(160.0 / h)) - 1                      # the sample at r = 160
        errs.append(abs(F[k] - exact[k]))
    assert 1.2 < errs[0] / errs[1] < 1.7 and 1.2 < errs[1] / errs[2] < 1.7
