"""Physical constants with the names ``mbproj2.physconstants`` exports.

Restated from memory of the public mbproj2 source (the package is neither vendored by the
reference nor installable here).  Only ``kpc_cm`` changes hot-path
numbers (Compton-y scaling, reference ``joxsz_funcs.py:459``; annulus volumes); the others
enter the sign-only mass veto (``joxsz_funcs.py:428-437``) or post-processing.
"""
kpc_cm = 3.0856776e21
kpc3_cm3 = kpc_cm ** 3
Mpc_cm = 3.0856776e24
Mpc_km = 3.0856776e19
km_cm = 1e5
keV_erg = 1.6022e-9
keV_K = 11.6048e6
boltzmann_erg_K = 1.3806503e-16
ne_nH = 1.2
mu_e = 1.17
mu_g = 1.6605e-24
solar_mass_g = 1.989e33
G_cgs = 6.67428e-8
yr_s = 31556926.0
