"""GPU evaluation of the component formulas at arbitrary radii (``jx_radial_profiles``).

Backs ``press_fun`` / ``press_derivative`` / ``vikhFunction`` / ``temp_fun`` / ``mass_fun`` of
``components.py`` (reference ``joxsz_funcs.py:275-301, 321-336, 375-395, 428-437``).  ``pars`` is the
reference's dict name -> object with ``.val``; a ``.val`` may be a scalar or an array of W walker
values.  Returns numpy, shaped like ``r_kpc`` for scalar parameters, else ``[W, *r_kpc.shape]``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_KIND_ARG = {"press": 0, "dpress": 1, "ne": 2, "tsz": 3, "tx": 4, "mass": 5}
# values for slots a formula does not read (they never influence the requested output)
_NEUTRAL = {"P_0": 1.0, "a": 1.0, "b": 1.0, "c": 0.0, "r_p": 1.0, "log(n_0)": 0.0, r"\beta": 1.0,
            "log(r_c)": 0.0, "log(r_s)": 0.0, r"\alpha": 0.0, r"\epsilon": 0.0, r"\gamma": 1.0,
            "log(n_{02})": 0.0, r"\beta_2": 1.0, "log(r_{c2})": 0.0, "log(T_X/T_{SZ})": 0.0,
            "Z": 0.0, "backscale": 1.0, "calibration": 1.0}


def evaluate(pars, r_kpc, kind, need, mode="single", mu_gas=0.61, device=None, per_walker_r=False):
    if not torch.cuda.is_available():
        raise _lib.JxError("profile evaluation runs on the GPU only (no CPU implementation in this package)")
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    vals, W, batched = {}, 1, False
    for name in need:
        v = np.asarray(pars[name].val, dtype=np.float64)
        if v.ndim > 0 and v.size > 1:
            v = v.reshape(-1)
            if batched and v.size != W:
                raise ValueError("walker-valued parameters have different lengths")
            W, batched = v.size, True
        else:
            v = v.reshape(-1)[:1]
        vals[name] = v
    full = np.empty((W, _lib.JX_NPAR))
    for i, name in enumerate(_lib.PARAM_SLOTS):
        full[:, i] = vals[name] if name in vals else _NEUTRAL[name]
    r = np.asarray(r_kpc, dtype=np.float64)
    if per_walker_r:        # r_kpc is [W, n]: walker w is evaluated on its own row
        if r.ndim != 2 or r.shape[0] != W:
            raise ValueError(f"per-walker radii must be [W={W}, n], got {r.shape}")
        n = r.shape[1]
        rf = np.ascontiguousarray(r.reshape(-1))
        r = r[0]
    else:
        rf = np.ascontiguousarray(r.reshape(-1))
        n = rf.size
    d_pars = torch.from_numpy(full).to(dev)
    d_r = torch.from_numpy(rf).to(dev)
    out = torch.empty((W, n), dtype=torch.float64, device=dev)
    args = [C.c_void_p(None)] * 6
    args[_KIND_ARG[kind]] = C.c_void_p(out.data_ptr())
    with torch.cuda.device(dev):
        rc = lib.jx_radial_profiles(C.c_void_p(d_pars.data_ptr()), W, 1 if mode == "double" else 0,
                                    C.c_void_p(d_r.data_ptr()), n, int(bool(per_walker_r)), float(mu_gas), *args,
                                    dev.index, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, None)
    res = out.cpu().numpy()
    if not batched:
        res = res[0].reshape(r.shape)
        return float(res) if res.ndim == 0 else res
    return res.reshape((W,) + r.shape)
