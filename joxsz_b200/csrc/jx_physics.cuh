// Radial profile formulas of the JoXSZ components, one walker's parameters at one radius.
// Reference: joxsz_funcs.py:275-301 (gNFW pressure and derivative), :375-395 (Vikhlinin density),
// :321-336 (temperatures), :428-437 (hydrostatic mass).  float64 throughout, like the reference.
#pragma once

#include "jx_common.cuh"

// mbproj2.physconstants.  Enter only the mass profile.
constexpr double JX_KPC_CM = 3.0856776e21;
constexpr double JX_KEV_ERG = 1.6022e-9;
constexpr double JX_MU_G = 1.6605e-24;
constexpr double JX_G_CGS = 6.67428e-8;
constexpr double JX_SOLAR_MASS_G = 1.989e33;

// Per-walker quantities that do not depend on radius, hoisted out of the radial loop.
//
// The radial formulas are evaluated in the log domain: with ln r tabulated once per grid and ln r_p,
// ln r_c, ln r_s taken once per walker, every power of the reference's expressions becomes part of one
// exponent, so a radius costs 4 exp + 3 log instead of 8 pow.  Differences from numpy's pow-by-pow
// evaluation are a few 1e-16 times the size of the exponent (<= ~30 here), far inside the 1e-12
// agreement the parity tests assert on the profiles.
struct jx_walker_pars {
    double P0, a, b, c, rp;
    double n0, n0sq, beta, rc, rs, alpha, eps, gamma;
    double n02sq, beta2, rc2;
    double tratio;           // 10**log(T_X/T_SZ)
    double e_press;          // (b-c)/a
    double e_core;           // 3 beta - alpha/2
    double e_outer;          // eps/gamma
    double ln_rp, ln_rc, ln_rs, ln_rc2;
    double inv_rc;           // 1 / r_c
    int dens_double;
};

JX_HD jx_walker_pars jx_prepare(const double* p, int dens_mode) {
    jx_walker_pars w;
    w.P0 = p[JX_P0]; w.a = p[JX_A]; w.b = p[JX_B]; w.c = p[JX_C]; w.rp = p[JX_RP];
    w.n0 = pow(10.0, p[JX_LOGN0]);
    w.n0sq = w.n0 * w.n0;
    w.beta = p[JX_BETA];
    w.rc = pow(10.0, p[JX_LOGRC]);
    w.rs = pow(10.0, p[JX_LOGRS]);
    w.alpha = p[JX_ALPHA]; w.eps = p[JX_EPS]; w.gamma = p[JX_GAMMA];
    w.dens_double = dens_mode;
    if (dens_mode) {
        double n02 = pow(10.0, p[JX_LOGN02]);
        w.n02sq = n02 * n02;
        w.beta2 = p[JX_BETA2];
        w.rc2 = pow(10.0, p[JX_LOGRC2]);
    } else {
        w.n02sq = 0.0; w.beta2 = 0.0; w.rc2 = 1.0;
    }
    w.tratio = pow(10.0, p[JX_LOGTRATIO]);
    w.e_press = (w.b - w.c) / w.a;
    w.e_core = 3.0 * w.beta - w.alpha / 2.0;
    w.e_outer = w.eps / w.gamma;
    w.ln_rp = log(w.rp); w.ln_rc = log(w.rc); w.ln_rs = log(w.rs); w.ln_rc2 = log(w.rc2);
    w.inv_rc = 1.0 / w.rc;
    return w;
}

// P(r) and dP/dr at radius r (lr = ln r).  With x = r/r_p:
//   P  = P0 / (x^c (1+x^a)^((b-c)/a))                                   (joxsz_funcs.py:287)
//   P' = -P0 (c + b x^a) / (r_p x^(c+1) (1+x^a)^((b-c+a)/a)) = -(c + b x^a) P / (r (1+x^a))   (:301)
JX_HD void jx_pressure(const jx_walker_pars& w, double r, double lr, double& press, double& dpress) {
    const double lx = lr - w.ln_rp;
    const double xa = exp(w.a * lx);
    const double opa = 1.0 + xa;
    press = w.P0 * exp(-(w.c * lx + w.e_press * log(opa)));
    dpress = -(w.c + w.b * xa) * press / (r * opa);
}

JX_HD double jx_pressure_only(const jx_walker_pars& w, double lr) {
    const double lx = lr - w.ln_rp;
    return w.P0 * exp(-(w.c * lx + w.e_press * log(1.0 + exp(w.a * lx))));
}

// n_e(r) (joxsz_funcs.py:389-395): sqrt(n0^2 x^-alpha / ((1+x^2)^(3 beta - alpha/2) (1+(r/rs)^gamma)^(eps/gamma)) [+ 2nd beta model])
JX_HD double jx_density(const jx_walker_pars& w, double r, double lr) {
    const double lxc = lr - w.ln_rc;
    const double xr = r * w.inv_rc, x2 = xr * xr;          // (r / r_c)^2 as the reference writes it: no exp
    const double t3 = exp(w.gamma * (lr - w.ln_rs));
    const double ex = -(w.alpha * lxc + w.e_core * log(1.0 + x2) + w.e_outer * log(1.0 + t3));
    if (!w.dens_double) return w.n0 * exp(0.5 * ex);
    const double y2 = exp(2.0 * (lr - w.ln_rc2));
    return sqrt(w.n0sq * exp(ex) + w.n02sq * exp(-3.0 * w.beta2 * log(1.0 + y2)));
}

// M(<r) in solar masses (joxsz_funcs.py:433-437)
JX_HD double jx_mass(double dpress_kpc, double ne, double r_kpc, double mu_gas) {
    double dpr_cm = dpress_kpc * JX_KEV_ERG / JX_KPC_CM;
    double r_cm = r_kpc * JX_KPC_CM;
    return -dpr_cm * (r_cm * r_cm) / (mu_gas * JX_MU_G * ne * JX_G_CGS) / JX_SOLAR_MASS_G;
}
