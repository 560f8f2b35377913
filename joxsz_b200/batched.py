"""BatchedLikelihood: the JoXSZ joint log-likelihood for every walker at once, on one B200.

``BatchedLikelihood(fit)(theta[W, ndim]) -> ll[W]`` is what an ensemble sampler calls once per
half-step (emcee ``vectorize=True`` convention) in place of ``W`` calls of the reference's
``getLikelihood`` (``joxsz_funcs.py:507-546``).  PyTorch only owns buffers and the stream; all
arithmetic happens in libjoxsz_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .packer import PackedSetup


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class BatchedLikelihood:
    def __init__(self, fit=None, packed: PackedSetup | None = None, max_walkers=1024, device=None, mode=None):
        if not torch.cuda.is_available():
            raise _lib.JxError("BatchedLikelihood needs a CUDA device: the likelihood has no CPU implementation")
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        if isinstance(device, torch.device):      # torch.device("cuda") has index None: that means the current device
            device = device.index if device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", int(device))
        if packed is None:
            packed = PackedSetup(fit, max_walkers=max_walkers, device=self.device.index)
        else:
            packed.device = self.device.index
            packed.max_walkers = max(int(max_walkers), 1) if max_walkers else packed.max_walkers
            packed._struct = None
        self.packed = packed
        self.ndim = packed.ndim
        self.max_walkers = packed.max_walkers
        handle = C.c_void_p()
        rc = self.lib.jx_create(C.byref(packed.struct()), C.byref(handle))
        _lib.check(rc, None)
        self._h = handle
        #: "staged" (default): the reference-shaped pipeline, every intermediate map exists (K1 K2 K3 K7 K5);
        #: "collapsed": the linear SZ chain folded into one constant operator, `ll` only (jx_loglike_collapsed).
        #: JX_MODE in the environment sets the default.
        import os
        self.mode = mode or os.environ.get("JX_MODE", "staged")
        if self.mode not in ("staged", "collapsed"):
            raise ValueError("mode must be 'staged' or 'collapsed'")
        self._pinned_in = None
        self._pinned_out = None
        self._dev_in = None
        self._dev_out = None

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "_h", None):
            self.lib.jx_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _theta_dev(self, theta):
        """[W, ndim] float64 CUDA tensor from numpy / torch input (host input goes through pinned memory)."""
        if isinstance(theta, torch.Tensor):
            t = theta
            if t.dim() == 1:
                t = t.unsqueeze(0)
            if t.device != self.device or t.dtype != torch.float64 or not t.is_contiguous():
                t = t.to(device=self.device, dtype=torch.float64).contiguous()
            return t
        a = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
        W = a.shape[0]
        if a.ndim != 2 or a.shape[1] != self.ndim:
            raise ValueError(f"theta must be [W, {self.ndim}], got {a.shape}")
        if self._pinned_in is None or self._pinned_in.shape[0] < W:
            self._pinned_in = torch.empty((max(W, 1), self.ndim), dtype=torch.float64).pin_memory()
        self._pinned_in[:W].copy_(torch.from_numpy(a))
        return self._pinned_in[:W].to(self.device, non_blocking=True)

    def _check_theta(self, t):
        if t.dim() != 2 or t.shape[1] != self.ndim:
            raise ValueError(f"theta must be [W, {self.ndim}], got {tuple(t.shape)}")
        if t.shape[0] > self.max_walkers:
            raise ValueError(f"W={t.shape[0]} exceeds max_walkers={self.max_walkers}")

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ hot path
    def loglike_device(self, theta_dev: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """theta_dev [W, ndim] CUDA float64 -> ll [W] CUDA float64, asynchronous on the current stream."""
        self._check_theta(theta_dev)
        W = theta_dev.shape[0]
        if out is None:
            out = self._new(W)
        if W == 0:              # an empty tensor has a NULL data pointer, which the C ABI rejects
            return out
        fn = self.lib.jx_loglike if self.mode == "staged" else self.lib.jx_loglike_collapsed
        with torch.cuda.device(self.device):
            rc = fn(self._h, _ptr(theta_dev), W, _ptr(out), self._stream())
        _lib.check(rc, self._h)
        return out

    #: host batches larger than this are fed to the device in chunks of this many walkers, so that staging chunk
    #: i+1 in pinned memory and its host->device copy overlap the kernels of chunk i (results do not depend on
    #: how a batch is split).  Measured on B200: chunks of 16 384 are 4 % slower than one 65 536-walker call (the
    #: small kernels and the launch count outweigh the hidden 0.7 ms of staging), hence the large value.
    HOST_CHUNK = 65536

    def _call_host(self, a, pinned_src=None):
        """numpy [W, ndim] -> numpy [W]: pinned staging, async H2D, kernels, async D2H, one synchronisation.
        ``pinned_src``: the same data as a pinned float64 torch tensor (the caller's own page-locked buffer): the
        staging copy is skipped and the host->device copy reads it directly."""
        W = a.shape[0]
        if a.shape[1] != self.ndim:
            raise ValueError(f"theta must be [W, {self.ndim}], got {a.shape}")
        if W > self.max_walkers:
            raise ValueError(f"W={W} exceeds max_walkers={self.max_walkers}")
        if self._pinned_in is None or self._pinned_in.shape[0] < W:
            self._pinned_in = torch.empty((max(W, 1), self.ndim), dtype=torch.float64).pin_memory()
        if self._pinned_out is None or self._pinned_out.shape[0] < W:
            self._pinned_out = torch.empty(max(W, 1), dtype=torch.float64).pin_memory()
        if self._dev_in is None or self._dev_in.shape[0] < W:
            self._dev_in = self._new(max(W, 1), self.ndim)
            self._dev_out = self._new(max(W, 1))
        src = torch.from_numpy(a) if pinned_src is None else pinned_src
        step = self.HOST_CHUNK if W > self.HOST_CHUNK else max(W, 1)
        for lo in range(0, W, step):
            hi = min(W, lo + step)
            if pinned_src is None:
                self._pinned_in[lo:hi].copy_(src[lo:hi])                   # host memcpy, overlaps the previous chunk
                self._dev_in[lo:hi].copy_(self._pinned_in[lo:hi], non_blocking=True)
            else:
                self._dev_in[lo:hi].copy_(src[lo:hi], non_blocking=True)
            self.loglike_device(self._dev_in[lo:hi], out=self._dev_out[lo:hi])
            self._pinned_out[lo:hi].copy_(self._dev_out[lo:hi], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._pinned_out[:W].numpy().copy()

    def __call__(self, theta):
        """numpy/torch [W, ndim] (or [ndim]) -> same kind of array [W] (or a float)."""
        was_torch = isinstance(theta, torch.Tensor)
        single = theta.dim() == 1 if was_torch else np.ndim(theta) == 1
        if was_torch and theta.is_cuda:
            ll = self.loglike_device(self._theta_dev(theta))
            return ll[0] if single else ll
        if (was_torch and theta.dim() == 2 and theta.dtype == torch.float64 and theta.is_contiguous()
                and theta.is_pinned()):
            res = self._call_host(theta.numpy(), pinned_src=theta)     # page-locked input: no staging copy
            return torch.from_numpy(res)
        a = theta.detach().numpy() if was_torch else np.asarray(theta)
        a = np.ascontiguousarray(np.atleast_2d(np.asarray(a, dtype=np.float64)))
        if a.ndim != 2:
            raise ValueError(f"theta must be [W, {self.ndim}], got {a.shape}")
        res = self._call_host(a)
        if was_torch:
            res = torch.from_numpy(res)
        return float(res[0]) if single else res

    # ------------------------------------------------------------------ parity taps
    def profiles(self, theta):
        """K1: dict(pp [W,nr], tsz [W,sep], ne_ann, tx_ann [W,na], flags [W], prior [W]) as numpy."""
        t = self._theta_dev(theta); self._check_theta(t)
        W, p = t.shape[0], self.packed
        o = dict(pp=self._new(W, p.nr), tsz=self._new(W, p.sep), ne_ann=self._new(W, p.na),
                 tx_ann=self._new(W, p.na), flags=self._new(W, dtype=torch.int32), prior=self._new(W))
        with torch.cuda.device(self.device):
            rc = self.lib.jx_profiles(self._h, _ptr(t), W, _ptr(o["pp"]), _ptr(o["tsz"]), _ptr(o["ne_ann"]),
                                      _ptr(o["tx_ann"]), _ptr(o["flags"]), _ptr(o["prior"]), self._stream())
        _lib.check(rc, self._h)
        return {k: v.cpu().numpy() for k, v in o.items()}

    def sz_project(self, theta):
        """K2: dict(y [W,nr], coef [W,4,nseg])."""
        t = self._theta_dev(theta); self._check_theta(t)
        W, p = t.shape[0], self.packed
        y, coef = self._new(W, p.nr), self._new(W, 4 * p.map_ops.nseg)
        with torch.cuda.device(self.device):
            rc = self.lib.jx_sz_project(self._h, _ptr(t), W, _ptr(y), _ptr(coef), self._stream())
        _lib.check(rc, self._h)
        return dict(y=y.cpu().numpy(), coef=coef.cpu().numpy().reshape(W, 4, -1))

    def sz_maps(self, theta, want=("y_2d", "conv_2d", "map_out")):
        """K3 full maps [W,N,N] (reference joxsz_funcs.py:462-467)."""
        t = self._theta_dev(theta); self._check_theta(t)
        W, N = t.shape[0], self.packed.N
        bufs = {k: (self._new(W, N, N) if k in want else None) for k in ("y_2d", "conv_2d", "map_out")}
        with torch.cuda.device(self.device):
            rc = self.lib.jx_sz_maps(self._h, _ptr(t), W, _ptr(bufs["y_2d"]), _ptr(bufs["conv_2d"]),
                                     _ptr(bufs["map_out"]), self._stream())
        _lib.check(rc, self._h)
        return {k: v.cpu().numpy() for k, v in bufs.items() if v is not None}

    def sz_profile(self, theta):
        """K3+K5: dict(row, bright [W,H], model [W,Nd], chisq [W], cint [W])."""
        t = self._theta_dev(theta); self._check_theta(t)
        W, p = t.shape[0], self.packed
        o = dict(row=self._new(W, p.H), bright=self._new(W, p.H), model=self._new(W, p.flux.size),
                 chisq=self._new(W), cint=self._new(W))
        with torch.cuda.device(self.device):
            rc = self.lib.jx_sz_profile(self._h, _ptr(t), W, _ptr(o["row"]), _ptr(o["bright"]), _ptr(o["model"]),
                                        _ptr(o["chisq"]), _ptr(o["cint"]), self._stream())
        _lib.check(rc, self._h)
        return {k: v.cpu().numpy() for k, v in o.items()}

    def xray(self, theta):
        """K4: dict(pred [W,nb,na], cash [W])."""
        t = self._theta_dev(theta); self._check_theta(t)
        W, p = t.shape[0], self.packed
        pred, cash = self._new(W, p.nb, p.na), self._new(W)
        with torch.cuda.device(self.device):
            rc = self.lib.jx_xray(self._h, _ptr(t), W, _ptr(pred), _ptr(cash), self._stream())
        _lib.check(rc, self._h)
        return dict(pred=pred.cpu().numpy(), cash=cash.cpu().numpy())

    def cash_from_profiles(self, pred):
        """``mylikeFromProfs`` on given predicted profiles [W, nb, na] -> [W] (numpy)."""
        p = self.packed
        a = np.ascontiguousarray(np.asarray(pred, dtype=np.float64)).reshape(-1, p.nb, p.na)
        if a.shape[0] > self.max_walkers:
            raise ValueError(f"W={a.shape[0]} exceeds max_walkers={self.max_walkers}")
        d = torch.from_numpy(a).to(self.device)
        out = self._new(a.shape[0])
        with torch.cuda.device(self.device):
            rc = self.lib.jx_cash_from_profiles(self._h, _ptr(d), a.shape[0], _ptr(out), self._stream())
        _lib.check(rc, self._h)
        return out.cpu().numpy()

    # ------------------------------------------------------------------ measurement
    def set_profiling(self, on: bool):
        _lib.check(self.lib.jx_set_profiling(self._h, int(bool(on))), self._h)

    def stage_times(self):
        """{stage: (total_ms, launches)} since the last call; resets the counters."""
        ms = (C.c_double * _lib.JX_NSTAGE)()
        n = (C.c_int64 * _lib.JX_NSTAGE)()
        _lib.check(self.lib.jx_stage_times(self._h, ms, n), self._h)
        return {name: (ms[i], n[i]) for i, name in enumerate(_lib.STAGE_NAMES)}
