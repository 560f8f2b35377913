"""ctypes binding of libjoxsz_b200.so (declarations mirror include/joxsz_b200.h one to one).

The library is the only compute path of this package.  If it is missing, :func:`load` raises --
there is no Python/numpy fallback for any likelihood stage.
"""
from __future__ import annotations

import ctypes as C
import os

JX_ABI_VERSION = 7
JX_NPAR = 19
JX_NSTAGE = 6
STAGE_NAMES = ("profiles", "project", "szmap", "xray", "tail", "filter")

# slot order of include/joxsz_b200.h `enum jx_param_slot`, keyed by the reference's parameter names
PARAM_SLOTS = (
    "P_0", "a", "b", "c", "r_p",
    "log(n_0)", r"\beta", "log(r_c)", "log(r_s)", r"\alpha", r"\epsilon", r"\gamma",
    "log(n_{02})", r"\beta_2", "log(r_{c2})",
    "log(T_X/T_{SZ})", "Z", "backscale", "calibration",
)
assert len(PARAM_SLOTS) == JX_NPAR

FLAG_PRIOR, FLAG_MASS, FLAG_RCRS, FLAG_XNONPOS = 1, 2, 4, 8

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class JxSetup(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("max_walkers", C.c_int32),
        ("ndim", C.c_int32), ("slot_src", C.c_int32 * JX_NPAR), ("slot_val", C.c_double * JX_NPAR),
        ("dens_mode", C.c_int32), ("exclude_unphy_mass", C.c_int32),
        ("prior_kind", _pi), ("prior_a", _pd), ("prior_b", _pd), ("prior_const", C.c_double),
        ("nr", C.c_int32), ("nt", C.c_int32), ("nmap", C.c_int32), ("nh", C.c_int32),
        ("npad", C.c_int32), ("nseg", C.c_int32),
        ("r_pp", _pd), ("proj_op", _pd), ("y_op", _pd), ("seg", _pi), ("dx", _pd), ("bhat", _pd),
        ("cmat", _pd), ("hf", _pd), ("dinv", _pd), ("filt_q", _pd), ("nbeam", C.c_int32), ("bmix", _pd),
        ("w_t0", _pd), ("nconv", C.c_int32), ("conv_T", _pd), ("conv_I", _pd),
        ("nd", C.c_int32), ("g_op", _pd), ("flux", _pd), ("flux_err", _pd),
        ("calc_integ", C.c_int32), ("w_integ", _pd), ("integ_mu", C.c_double), ("integ_sig", C.c_double),
        ("na", C.c_int32), ("nb", C.c_int32), ("ntab", C.c_int32),
        ("midpt_kpc", _pd), ("projvols", _pd), ("tlog", _pd),
        ("tmin", C.c_double), ("tmax", C.c_double),
        ("lnrate0", _pd), ("lnrate1", _pd), ("cts", _pd), ("srcscale", _pd), ("bkgterm", _pd),
    ]


# name -> (restype, argtypes); every symbol include/joxsz_b200.h declares
_vp = C.c_void_p
PROTOTYPES = {
    "jx_create": (C.c_int, [C.POINTER(JxSetup), C.POINTER(_vp)]),
    "jx_destroy": (None, [_vp]),
    "jx_last_error": (C.c_char_p, [_vp]),
    "jx_loglike": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp]),
    "jx_loglike_collapsed": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp]),
    "jx_profiles": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jx_sz_project": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp]),
    "jx_sz_maps": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "jx_sz_profile": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "jx_xray": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp]),
    "jx_cash_from_profiles": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp]),
    "jx_radial_profiles": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_double,
                                     _vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, _vp]),
    "jx_stretch_propose": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                     C.c_uint64, C.c_uint64, _vp, _vp, _vp, C.c_int32, _vp]),
    "jx_stretch_accept": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp, C.c_int32, _vp]),
    "jx_stretch_permutation": (C.c_int, [_vp, C.c_int32, C.c_uint64, C.c_uint64, _vp, _vp, C.POINTER(C.c_size_t),
                                         C.c_int32, _vp]),
    "jx_stretch_advance": (C.c_int, [_vp, C.c_uint64, C.c_int32, _vp]),
    "jx_stretch_scatter": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_int32,
                                     C.c_int32, _vp]),
    "jx_stretch_accept_p2p": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp,
                                        _vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp,
                                        C.c_int32, _vp]),
    "jx_stretch_scatter_p2p": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _vp, C.c_int32, _vp]),
    "jx_set_profiling": (C.c_int, [_vp, C.c_int32]),
    "jx_stage_times": (C.c_int, [_vp, _pd, C.POINTER(C.c_int64)]),
    "jx_build_info": (C.c_char_p, []),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libjoxsz_b200.so")
_lib = None


class JxError(RuntimeError):
    """A libjoxsz_b200 call returned a negative status."""


def load():
    """dlopen the in-tree library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JxError(f"{LIB_PATH} is missing: build it with `python -m joxsz_b200.build` "
                      "(there is no CPU implementation to fall back to)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc == 0:
        return
    lib = load()
    msg = lib.jx_last_error(handle)
    raise JxError(f"libjoxsz_b200 status {rc}: {(msg or b'').decode('utf-8', 'replace')}")
