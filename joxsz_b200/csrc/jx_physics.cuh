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
struct jx_walker_pars {
    double P0, a, b, c, rp;
    double n0sq, beta, rc, rs, alpha, eps, gamma;
    double n02sq, beta2, rc2;
    double tratio;           // 10**log(T_X/T_SZ)
    double e_press;          // (b-c)/a
    double e_dpress;         // (b-c+a)/a
    double e_core;           // 3 beta - alpha/2
    double e_outer;          // eps/gamma
    int dens_double;
};

JX_HD jx_walker_pars jx_prepare(const double* p, int dens_mode) {
    jx_walker_pars w;
    w.P0 = p[JX_P0]; w.a = p[JX_A]; w.b = p[JX_B]; w.c = p[JX_C]; w.rp = p[JX_RP];
    double n0 = pow(10.0, p[JX_LOGN0]);
    w.n0sq = n0 * n0;
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
    w.e_dpress = (w.b - w.c + w.a) / w.a;
    w.e_core = 3.0 * w.beta - w.alpha / 2.0;
    w.e_outer = w.eps / w.gamma;
    return w;
}

// P(r) and dP/dr.  x^c, x^a and (1+x^a) are shared between the two, exactly the factors the
// reference raises to powers (joxsz_funcs.py:287, 301).
JX_HD void jx_pressure(const jx_walker_pars& w, double r, double& press, double& dpress) {
    double x = r / w.rp;
    double xc = pow(x, w.c);
    double xa = pow(x, w.a);
    double opa = 1.0 + xa;
    press = w.P0 / (xc * pow(opa, w.e_press));
    dpress = -w.P0 * (w.c + w.b * xa) / (w.rp * pow(x, w.c + 1.0) * pow(opa, w.e_dpress));
}

JX_HD double jx_pressure_only(const jx_walker_pars& w, double r) {
    double x = r / w.rp;
    return w.P0 / (pow(x, w.c) * pow(1.0 + pow(x, w.a), w.e_press));
}

// n_e(r) (joxsz_funcs.py:389-395)
JX_HD double jx_density(const jx_walker_pars& w, double r) {
    double x = r / w.rc;
    double res = w.n0sq * pow(x, -w.alpha)
                 / (pow(1.0 + x * x, w.e_core) * pow(1.0 + pow(r / w.rs, w.gamma), w.e_outer));
    if (w.dens_double) {
        double x2 = r / w.rc2;
        res += w.n02sq / pow(1.0 + x2 * x2, 3.0 * w.beta2);
    }
    return sqrt(res);
}

// M(<r) in solar masses (joxsz_funcs.py:433-437)
JX_HD double jx_mass(double dpress_kpc, double ne, double r_kpc, double mu_gas) {
    double dpr_cm = dpress_kpc * JX_KEV_ERG / JX_KPC_CM;
    double r_cm = r_kpc * JX_KPC_CM;
    return -dpr_cm * (r_cm * r_cm) / (mu_gas * JX_MU_G * ne * JX_G_CGS) / JX_SOLAR_MASS_G;
}
