"""Device-resident ensemble sampler: emcee's stretch move driven by the batched likelihood.

Replaces, for the path this repo accelerates, ``emcee.EnsembleSampler`` as the reference uses it
(``joxsz_main.py:203-210``; iteration schedule ``joxsz_funcs.py:548-635``): instead of ``pool.map`` of
one ``getLikelihood`` call per walker, each half-step proposes, evaluates and accepts every active
walker at once on the GPU (``jx_stretch_propose`` -> ``jx_loglike`` -> ``jx_stretch_accept`` ->
all-gather -> ``jx_stretch_scatter``).

Multi-GPU: one process per GPU.  Every rank holds the whole ensemble (``coords [W, ndim]``, ``lp [W]``,
kept identical by construction); in a half-step rank ``g`` handles the contiguous slice
``[g*per, (g+1)*per)`` of the active colour and the only collective is one all-gather of the packed
results ``[per, ndim + 2]`` (new position, new log-prob, accepted flag) per half-step.  Random numbers
are counter based (Philox keyed by seed, walker, iteration), so the chain does not depend on the number
of ranks.

The sampler itself does no arithmetic: proposals, acceptance and scatter are CUDA kernels behind the
C ABI (``ops=None``).  The ``ops`` hook exists so that the index/sharding logic can be exercised on CPU
with a numpy model of those kernels (tests only); the product path raises without the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class CudaStretchOps:
    """``jx_stretch_*`` through ctypes (include/joxsz_b200.h)."""

    launches_per_half_step = 3
    launches_per_permutation = 5        # key kernel + CUB radix sort passes (64-bit keys, onesweep)

    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise _lib.JxError("the stretch-move kernels are CUDA only (no CPU implementation)")
        self.lib = _lib.load()
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self._ws = {}
        #: when set (a [1] int64 CUDA tensor), every kernel adds ``*iter_dev`` to the ``iteration`` it is given:
        #: the sampler captures an iteration in a CUDA graph with ``iteration`` = 0 / 1 and the counter on the device
        self.iter_dev = None

    def _iter_ptr(self):
        return _ptr(self.iter_dev)

    def advance(self, by=1):
        _lib.check(self.lib.jx_stretch_advance(_ptr(self.iter_dev), C.c_uint64(by), self.index, self._stream()))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def permutation(self, perm, seed, iteration):
        """Colouring permutation of this iteration, generated on the device (identical on every rank)."""
        nall = perm.shape[0]
        # one sort workspace per destination buffer: the permutations of two consecutive iterations are produced
        # on different streams (and, in graph mode, by different graphs) and may overlap in time
        key = (nall, perm.data_ptr())
        ws = self._ws.get(key)
        if ws is None:
            need = C.c_size_t(0)
            _lib.check(self.lib.jx_stretch_permutation(None, nall, C.c_uint64(0), C.c_uint64(0), None, None,
                                                       C.byref(need), self.index, None))
            ws = self._ws[key] = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        nbytes = C.c_size_t(ws.numel())
        rc = self.lib.jx_stretch_permutation(_ptr(perm), nall, C.c_uint64(seed), C.c_uint64(iteration),
                                             self._iter_ptr(), _ptr(ws), C.byref(nbytes), self.index, self._stream())
        _lib.check(rc)

    def propose(self, coords, perm, split, r_first, r_count, a, seed, iteration, prop, factor):
        nall, ndim = coords.shape
        rc = self.lib.jx_stretch_propose(_ptr(coords), _ptr(perm), nall, ndim, split, r_first, r_count, float(a),
                                         C.c_uint64(seed), C.c_uint64(iteration), self._iter_ptr(), _ptr(prop),
                                         _ptr(factor), self.index, self._stream())
        _lib.check(rc)

    def accept(self, coords, lp, perm, split, r_first, r_count, prop, lp_new, factor, seed, iteration, packed):
        nall, ndim = coords.shape
        rc = self.lib.jx_stretch_accept(_ptr(coords), _ptr(lp), _ptr(perm), nall, ndim, split, r_first, r_count,
                                        _ptr(prop), _ptr(lp_new), _ptr(factor), C.c_uint64(seed),
                                        C.c_uint64(iteration), self._iter_ptr(), _ptr(packed), self.index,
                                        self._stream())
        _lib.check(rc)

    def scatter(self, coords, lp, naccept, perm, split, packed_all, ns):
        nall, ndim = coords.shape
        rc = self.lib.jx_stretch_scatter(_ptr(coords), _ptr(lp), _ptr(naccept), _ptr(perm), nall, ndim, split,
                                         _ptr(packed_all), ns, self.index, self._stream())
        _lib.check(rc)

    # ---- accept + exchange over peer memory (no collective launch)
    def accept_p2p(self, coords, lp, perm, split, r_first, r_count, prop, lp_new, factor, seed, iteration, px):
        nall, ndim = coords.shape
        rc = self.lib.jx_stretch_accept_p2p(_ptr(coords), _ptr(lp), _ptr(perm), nall, ndim, split, r_first, r_count,
                                            _ptr(prop), _ptr(lp_new), _ptr(factor), C.c_uint64(seed),
                                            C.c_uint64(iteration), self._iter_ptr(), px.peer_packed, px.peer_flags,
                                            px.world, px.rank, px.per0, _ptr(px.done), self.index, self._stream())
        _lib.check(rc)

    def scatter_p2p(self, coords, lp, naccept, perm, split, ns, per, iteration, px):
        nall, ndim = coords.shape
        rc = self.lib.jx_stretch_scatter_p2p(_ptr(coords), _ptr(lp), _ptr(naccept), _ptr(perm), nall, ndim, split,
                                             _ptr(px.packed), _ptr(px.flags), ns, px.world, per, px.per0,
                                             C.c_uint64(iteration), self._iter_ptr(), self.index, self._stream())
        _lib.check(rc)


class PeerExchange:
    """Buffers of the fused accept + exchange: ``packed`` [2, world * per0, ndim + 2] float64 and ``flags`` [2, world]
    int64 in torch symmetric memory (every rank of the node can store into every rank's copy over NVLink), and the
    host arrays of the peers' device addresses the kernels take."""

    def __init__(self, world, rank, per0, ndim, device, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank, self.per0 = world, rank, per0
        self.packed = symm.empty((2 * world * per0 * (ndim + 2),), dtype=torch.float64, device=device)
        self.flags = symm.empty((2 * world,), dtype=torch.int64, device=device)
        self.packed.zero_()
        self.flags.zero_()
        self.done = torch.zeros((1,), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        group = group if group is not None else dist.group.WORLD
        hp = symm.rendezvous(self.packed, group)
        hf = symm.rendezvous(self.flags, group)
        self._handles = (hp, hf)
        self.peer_packed = (C.c_uint64 * world)(*[int(p) for p in hp.buffer_ptrs])
        self.peer_flags = (C.c_uint64 * world)(*[int(p) for p in hf.buffer_ptrs])
        if int(hp.buffer_ptrs[rank]) != self.packed.data_ptr() or int(hf.buffer_ptrs[rank]) != self.flags.data_ptr():
            raise RuntimeError("symmetric memory: local buffer address mismatch")
        dist.barrier(group=group)             # every rank's flags are zero before anyone stores into them
        torch.cuda.synchronize(device)


def shard_bounds(ns: int, world: int, rank: int):
    """Slice of the ``ns`` active walkers of a half-step handled by ``rank``: (per, r_first, r_count).
    ``per`` = ceil(ns / world) is the static row count every rank contributes to the all-gather."""
    per = (ns + world - 1) // world
    r_first = min(rank * per, ns)
    r_count = max(0, min(per, ns - r_first))
    return per, r_first, r_count


class State:
    """What one iteration of :meth:`EnsembleSampler.sample` yields (emcee ``State`` look-alike);
    arrays are copied to the host only when read."""

    def __init__(self, sampler):
        self._s = sampler
        self.random_state = None
        self.blobs = None

    @property
    def coords(self):
        return self._s.coords_host()

    @property
    def log_prob(self):
        return self._s.log_prob_host()

    def __iter__(self):          # emcee-2 style unpacking: pos, lnprob, rstate = result[:3]
        return iter((self.coords, self.log_prob, self.random_state))

    def __getitem__(self, i):
        return (self.coords, self.log_prob, self.random_state)[i]


class EnsembleSampler:
    """Subset of ``emcee.EnsembleSampler`` the reference touches, on device.

    ``log_prob_fn``: a :class:`joxsz_b200.batched.BatchedLikelihood` (anything with
    ``loglike_device(theta[W, ndim]) -> ll[W]``), or the bound ``fit.getLikelihood`` exactly as
    ``joxsz_main.py:206`` passes it (its engine is looked up with :func:`joxsz_b200.funcs.engine_for`).
    ``pool`` / ``backend`` are accepted and ignored (walkers are batched on the GPU; chains live in
    memory, see :meth:`save_npz`).
    """

    def __init__(self, nwalkers, ndim, log_prob_fn, pool=None, backend=None, a=2.0, seed=None, world_size=1,
                 rank=0, group=None, ops=None, device=None, vectorize=True, moves=None, graph=None, chain_shard=False, exchange="auto"):
        if nwalkers < 2 * ndim:
            raise ValueError("The number of walkers needs to be at least twice the dimension of the problem")
        if moves is not None:
            raise ValueError("only the default StretchMove(a) is implemented")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.world, self.rank, self.group = int(world_size), int(rank), group
        if seed is None:
            seed = int(np.random.randint(0, 2**31 - 1))      # np.random.seed(seed) upstream (joxsz_main.py:204)
        self.seed = int(seed)
        self.per0 = shard_bounds((self.nwalkers + 1) // 2, self.world, self.rank)[0]
        need = max(self.per0, shard_bounds(self.nwalkers, self.world, self.rank)[0])
        self.fit = None
        if hasattr(log_prob_fn, "loglike_device"):
            self.engine = log_prob_fn
        else:
            fit = getattr(log_prob_fn, "__self__", None)
            if fit is None or not hasattr(fit, "thawed"):
                raise TypeError("log_prob_fn must be a BatchedLikelihood or the bound fit.getLikelihood")
            from .funcs import engine_for
            self.fit = fit
            self.engine = engine_for(fit, need)
        if self.engine.ndim != self.ndim:
            raise ValueError(f"ndim={ndim} but the likelihood has {self.engine.ndim} thawed parameters")
        if getattr(self.engine, "max_walkers", need) < need:
            raise ValueError(f"likelihood engine capacity {self.engine.max_walkers} < {need} walkers per call")
        self.device = torch.device(device) if device is not None else getattr(self.engine, "device", None)
        if self.device is None:
            raise ValueError("device unknown")
        self.ops = ops if ops is not None else CudaStretchOps(self.device)
        #: one ensemble iteration as ONE CUDA graph launch (the ~20 kernel launches, the side-stream fork/joins and the
        #: all-gathers of an iteration are replayed by the driver instead of being issued from Python one by one).
        #: None = on whenever the kernels are the CUDA ones; the first GRAPH_EAGER_STEPS iterations always run eagerly
        #: (they warm up NCCL and the allocator), then the iteration is captured once and replayed.
        self.use_graph = isinstance(self.ops, CudaStretchOps) if graph is None else bool(graph)
        if self.use_graph and not isinstance(self.ops, CudaStretchOps):
            raise ValueError("graph=True needs the CUDA stretch-move kernels")
        #: True: a rank keeps only the chain of its own walkers [rank*W/world, (rank+1)*W/world) -- every rank holds the
        #: whole ensemble state, but the stored chain of a long many-walker run (nsteps x W x ndim doubles) need not be
        #: replicated on every GPU; get_chain() / get_log_prob() then return the local shard
        self.chain_shard = bool(chain_shard) and self.world > 1
        lo = (self.nwalkers * self.rank) // self.world if self.chain_shard else 0
        hi = (self.nwalkers * (self.rank + 1)) // self.world if self.chain_shard else self.nwalkers
        self.chain_walkers = (lo, hi)
        #: how the half-step results travel between the ranks: "p2p" = the accept kernel stores its rows into every rank's
        #: buffer over NVLink peer memory and the scatter kernel waits on flags (no collective launch); "nccl" = accept,
        #: then one all_gather_into_tensor; "auto" = p2p when the CUDA kernels run on > 1 rank and torch symmetric
        #: memory can be set up, else nccl
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'nccl'")
        self.exchange = exchange
        self._px = None
        self._px_failed = None
        self._graph = None
        self._graph_failed = None
        self.eager_only = False               # True: launch kernel by kernel even if a graph exists (per-kernel timers)
        self.initspread = 0.1
        self.pos0 = None
        self.backend = self                   # mcmc.backend.get_chain()/get_log_prob()/reset(...)
        self.iteration = 0                    # global Philox counter: never reset, so restarts do not repeat draws
        self._steps_total = 0                 # steps since construction (reset() does not clear it)
        self.aux_launches = 0
        self._coords = self._lp = None
        self._chain = self._chain_lp = None
        self._nstored = 0
        self._alloc()

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        dev, f64 = self.device, torch.float64
        W, nd, per = self.nwalkers, self.ndim, self.per0
        self._coords = torch.zeros((W, nd), dtype=f64, device=dev)
        self._lp = torch.full((W,), -math.inf, dtype=f64, device=dev)
        self._naccept = torch.zeros((W,), dtype=torch.int32, device=dev)
        self._prop = torch.zeros((per, nd), dtype=f64, device=dev)
        self._factor = torch.zeros((per,), dtype=f64, device=dev)
        self._lpnew = torch.zeros((per,), dtype=f64, device=dev)
        self._packed = torch.zeros((per, nd + 2), dtype=f64, device=dev)
        self._packed_all = self._packed if self.world == 1 else torch.zeros((per * self.world, nd + 2), dtype=f64,
                                                                            device=dev)
        # the colouring permutation of iteration i depends on (seed, i) only: two buffers, the one of the next
        # iteration is produced on a side stream while this iteration's kernels run (CUDA ops only)
        self._perm2 = [torch.zeros((W,), dtype=torch.int32, device=dev) for _ in range(2)]
        self._perm_iter = [None, None]
        self._perm = self._perm2[0]
        self._side = self._perm_ready = self._step_done = None
        if dev.type == "cuda" and isinstance(self.ops, CudaStretchOps):
            self._side = torch.cuda.Stream(device=dev)
            self._perm_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._step_done = torch.cuda.Event()
        self._steps_done = 0
        # graph mode: iteration counter of the Philox streams in device memory, the permutation the graph reads and the
        # one its side branch produces for the next iteration, re-pack buffer of odd ensembles
        self._iter_dev = self._perm_g = self._perm_next = None
        self._iter_dev_value = None
        self._pa_buf = None
        if self.world > 1:
            self._pa_buf = torch.zeros((per * self.world, nd + 2), dtype=f64, device=dev)
        if self.world > 1 and self.exchange != "nccl" and isinstance(self.ops, CudaStretchOps):
            try:
                self._px = PeerExchange(self.world, self.rank, per, nd, dev, self.group)
            except Exception as e:
                self._px_failed = f"{type(e).__name__}: {e}"
                if self.exchange == "p2p":
                    raise
        if self.world > 1 and isinstance(self.ops, CudaStretchOps):
            # every rank must take the same path: p2p only if every rank could set it up
            import torch.distributed as dist
            ok = torch.tensor([1 if self._px is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self._px = None
        if self.use_graph:
            self._iter_dev = torch.zeros((1,), dtype=torch.int64, device=dev)
            self._perm_g = torch.zeros((W,), dtype=torch.int32, device=dev)
            self._perm_next = torch.zeros((W,), dtype=torch.int32, device=dev)

    def _all_gather(self, out, inp):
        import torch.distributed as dist
        try:
            dist.all_gather_into_tensor(out, inp, group=self.group)
        except (RuntimeError, NotImplementedError):      # backends without the flat variant
            parts = list(out.view(self.world, *inp.shape).unbind(0))
            dist.all_gather(parts, inp, group=self.group)

    # ------------------------------------------------------------------ state
    def initialize(self, p0, log_prob=None):
        """Load the starting ensemble [W, ndim]; log-probs are evaluated (sharded over ranks) if absent."""
        p0 = np.ascontiguousarray(np.asarray(p0, dtype=np.float64))
        if p0.shape != (self.nwalkers, self.ndim):
            raise ValueError(f"incompatible input dimensions {p0.shape}")
        if not np.all(np.isfinite(p0)):
            raise ValueError("At least one parameter value was infinite or NaN")
        self._coords.copy_(torch.from_numpy(p0))
        if log_prob is not None:
            self._lp.copy_(torch.from_numpy(np.asarray(log_prob, dtype=np.float64)))
        else:
            per, first, count = shard_bounds(self.nwalkers, self.world, self.rank)
            mine = torch.full((per,), -math.inf, dtype=torch.float64, device=self.device)
            if count:
                mine[:count] = self.engine.loglike_device(self._coords[first:first + count].contiguous())
            if self.world == 1:
                self._lp.copy_(mine[:self.nwalkers])
            else:
                full = torch.empty((per * self.world,), dtype=torch.float64, device=self.device)
                self._all_gather(full, mine)
                self._lp.copy_(full[:self.nwalkers])
        if bool(torch.isnan(self._lp).any()):
            raise ValueError("Probability function returned NaN")
        self._naccept.zero_()
        self._steps_done = 0

    def coords_host(self):
        return self._coords.cpu().numpy()

    def log_prob_host(self):
        return self._lp.cpu().numpy()

    def evals_per_rank_per_launch(self):
        return self.per0

    # ------------------------------------------------------------------ one ensemble iteration
    GRAPH_EAGER_STEPS = 2

    def _half_steps(self, perm, it):
        """The two half-steps of an iteration on the current stream (``it``: the iteration the kernels are given; in
        graph mode they add the device-side counter to it)."""
        W = self.nwalkers
        for split in (0, 1):
            ns = (W - split + 1) // 2
            per, first, count = shard_bounds(ns, self.world, self.rank)
            if count:
                prop = self._prop[:count]
                self.ops.propose(self._coords, perm, split, first, count, self.a, self.seed, it, prop, self._factor)
                self.engine.loglike_device(prop, out=self._lpnew[:count])
            if self._px is not None:
                # accept + exchange in one kernel: rows go straight into every rank's buffer over NVLink
                self.ops.accept_p2p(self._coords, self._lp, perm, split, first, count, self._prop, self._lpnew,
                                    self._factor, self.seed, it, self._px)
                self.ops.scatter_p2p(self._coords, self._lp, self._naccept, perm, split, ns, per, it, self._px)
                self.aux_launches += 2
                continue
            if count:
                self.ops.accept(self._coords, self._lp, perm, split, first, count, prop, self._lpnew,
                                self._factor, self.seed, it, self._packed)
            if self.world > 1:
                # static shape [per0, ndim+2] per rank; rows beyond `ns` are never read by scatter
                self._all_gather(self._packed_all, self._packed)
                if per != self.per0:       # odd ensembles: the second colour has one walker fewer -> re-pack rows
                    pa = self._pa_buf[:self.world * per]
                    pa.view(self.world, per, -1).copy_(self._packed_all.view(self.world, self.per0, -1)[:, :per])
                else:
                    pa = self._packed_all
            else:
                pa = self._packed
            self.ops.scatter(self._coords, self._lp, self._naccept, perm, split, pa, ns)
            self.aux_launches += self.ops.launches_per_half_step

    def _step_eager(self):
        it = self.iteration
        cur = it & 1
        self._perm = self._perm2[cur]
        if self._perm_iter[cur] != it:               # first step (or a jump of the counter): produce it in line
            self.ops.permutation(self._perm, self.seed, it)
            self._perm_iter[cur] = it
        elif self._side is not None:
            torch.cuda.current_stream(self.device).wait_event(self._perm_ready[cur])
        self.aux_launches += getattr(self.ops, "launches_per_permutation", 0)
        if self._side is not None:
            # next iteration's permutation on the side stream; its buffer was last read by the previous step
            nxt = cur ^ 1
            self._side.wait_event(self._step_done) if self._steps_done else self._side.wait_stream(
                torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._side):
                self.ops.permutation(self._perm2[nxt], self.seed, it + 1)
                self._perm_ready[nxt].record(self._side)
            self._perm_iter[nxt] = it + 1
        self._half_steps(self._perm, it)
        if self._side is not None:
            self._step_done.record(torch.cuda.current_stream(self.device))

    def _capture(self):
        """Record one iteration as a CUDA graph.  Every kernel takes its Philox iteration as ``k + *iter_dev`` with the
        counter in device memory, so the same graph is replayed for every iteration: the side branch sorts the
        colouring permutation of iteration ``*iter_dev + 1`` while the main branch runs the two half-steps on the
        current one; the tail copies it over and bumps the counter."""
        ops, dev = self.ops, self.device
        main = torch.cuda.current_stream(dev)
        ops.iter_dev = self._iter_dev
        try:
            self._sync_graph_state()
            ops.permutation(self._perm_next, self.seed, 1)      # creates the sort workspace outside the capture
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            launches0 = self.aux_launches
            with torch.cuda.graph(g):
                cap = torch.cuda.current_stream(dev)
                self._side.wait_stream(cap)
                with torch.cuda.stream(self._side):
                    ops.permutation(self._perm_next, self.seed, 1)
                self._half_steps(self._perm_g, 0)
                cap.wait_stream(self._side)
                self._perm_g.copy_(self._perm_next)
                ops.advance(1)
            self.aux_launches = launches0
            self._graph = g
        finally:
            ops.iter_dev = None
        del main

    def _sync_graph_state(self):
        """Device-side counter and current permutation = host-side ``self.iteration`` (first replay, or after the
        counter was moved from outside)."""
        ops = self.ops
        self._iter_dev.fill_(self.iteration)
        keep, ops.iter_dev = ops.iter_dev, self._iter_dev
        try:
            ops.permutation(self._perm_g, self.seed, 0)
        finally:
            ops.iter_dev = keep
        self._iter_dev_value = self.iteration

    def step(self):
        """One stretch-move iteration: two half-steps, every walker proposed and evaluated once."""
        if (self.use_graph and not self.eager_only and self._graph_failed is None
                and self._steps_total >= self.GRAPH_EAGER_STEPS):
            if self._graph is None:
                try:
                    self._capture()
                except Exception as e:      # e.g. a collective backend that cannot be captured: stay eager, say so
                    self._graph_failed = f"{type(e).__name__}: {e}"
                    self._graph = None
                    self.ops.iter_dev = None
                    torch.cuda.synchronize(self.device)
            if self._graph is not None:
                if self._iter_dev_value != self.iteration:
                    self._sync_graph_state()
                self._graph.replay()
                self._perm = self._perm_g
                self.aux_launches += (2 * self.ops.launches_per_half_step + self.ops.launches_per_permutation + 2)
                self.iteration += 1
                self._iter_dev_value = self.iteration
                self._steps_done += 1
                self._steps_total += 1
                return
        self._step_eager()
        self.iteration += 1
        self._steps_done += 1
        self._steps_total += 1

    @property
    def graph_active(self):
        return self._graph is not None

    def release_graph(self):
        """Drop the captured iteration (it is re-captured on the next step).  Call it before
        ``torch.distributed.destroy_process_group()``: a live CUDA graph that holds captured NCCL collectives keeps the
        communicator busy and the teardown waits for it forever."""
        if self._graph is not None:
            torch.cuda.synchronize(self.device)
            self._graph = None
            self._iter_dev_value = None

    def close(self):
        """Release the graph and the peer-memory buffers (end of a multi-process run)."""
        self.release_graph()
        self._px = None

    def launches_per_step(self, kernels_per_loglike=6):
        """Kernels of this library launched (or replayed) per ensemble iteration on this rank."""
        n = 2 * (kernels_per_loglike + self.ops.launches_per_half_step) + getattr(self.ops, "launches_per_permutation", 0)
        return n + (1 if self.graph_active else 0)              # + the counter kernel of the graph's tail

    # ------------------------------------------------------------------ emcee-like driver API
    def reset(self, nwalkers=None, ndim=None):
        """``backend.reset(nwalkers, ndim)`` / ``sampler.reset()``: drop the stored chain."""
        if nwalkers is not None and (int(nwalkers), int(ndim)) != (self.nwalkers, self.ndim):
            raise ValueError("cannot change the ensemble shape of a device sampler")
        self._chain = self._chain_lp = None
        self._nstored = 0
        self._naccept.zero_()
        self._steps_done = 0

    def sample(self, initial_state, iterations=1, thin_by=1, thin=None, store=True, storechain=None, progress=False,
               log_prob0=None):
        """Generator over iterations, emcee 3 conventions: ``thin_by=k`` runs ``iterations*k`` steps and
        stores/yields every k-th; the deprecated ``thin=k`` runs ``iterations`` steps, stores every k-th,
        yields every step (this is the form the reference uses, ``joxsz_funcs.py:593-622``)."""
        if storechain is not None:
            store = storechain
        if isinstance(initial_state, State):
            initial_state = initial_state.coords
        self.initialize(np.asarray(initial_state), log_prob0)
        if thin is not None:
            thin = int(thin)
            if thin <= 0:
                raise ValueError("Invalid thinning argument")
            yield_step, checkpoint_step, total = 1, thin, int(iterations)
            nsave = int(iterations) // thin
        else:
            thin_by = int(thin_by)
            if thin_by <= 0:
                raise ValueError("Invalid thinning argument")
            yield_step = checkpoint_step = thin_by
            total = int(iterations) * thin_by
            nsave = int(iterations)
        if store:
            self._grow(nsave)
        bar = None
        if progress:
            try:
                from tqdm import tqdm
                bar = tqdm(total=total)
            except Exception:
                bar = None
        state = State(self)
        for i in range(1, total + 1):
            self.step()
            if store and i % checkpoint_step == 0:
                lo, hi = self.chain_walkers
                self._chain[self._nstored].copy_(self._coords[lo:hi])
                self._chain_lp[self._nstored].copy_(self._lp[lo:hi])
                self._nstored += 1
            if bar is not None:
                bar.update(1)
            if i % yield_step == 0:
                yield state
        if bar is not None:
            bar.close()

    def run_mcmc(self, initial_state, nsteps, **kw):
        state = None
        for state in self.sample(initial_state, iterations=nsteps, **kw):
            pass
        return state

    def _grow(self, nsave):
        need = self._nstored + nsave
        nw = self.chain_walkers[1] - self.chain_walkers[0]
        if self._chain is None:
            self._chain = torch.empty((need, nw, self.ndim), dtype=torch.float64, device=self.device)
            self._chain_lp = torch.empty((need, nw), dtype=torch.float64, device=self.device)
        elif self._chain.shape[0] < need:
            c = torch.empty((need, nw, self.ndim), dtype=torch.float64, device=self.device)
            l = torch.empty((need, nw), dtype=torch.float64, device=self.device)
            c[:self._nstored] = self._chain[:self._nstored]
            l[:self._nstored] = self._chain_lp[:self._nstored]
            self._chain, self._chain_lp = c, l

    def get_chain(self, flat=False, thin=1, discard=0):
        if self._chain is None or self._nstored == 0:
            raise AttributeError("you must run the sampler with 'store == True' before accessing the results")
        v = self._chain[discard:self._nstored:thin].cpu().numpy()
        return v.reshape(-1, self.ndim) if flat else v

    def get_log_prob(self, flat=False, thin=1, discard=0):
        if self._chain_lp is None or self._nstored == 0:
            raise AttributeError("you must run the sampler with 'store == True' before accessing the results")
        v = self._chain_lp[discard:self._nstored:thin].cpu().numpy()
        return v.reshape(-1) if flat else v

    @property
    def chain(self):
        """[nwalkers, nsteps, ndim] (emcee's legacy layout, read at ``joxsz_main.py:213``)."""
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def acceptance_fraction(self):
        return self._naccept.cpu().numpy() / max(self._steps_done, 1)

    def mean_acceptance(self):
        return float(self._naccept.double().mean().item()) / max(self._steps_done, 1)

    def save_npz(self, path, **attrs):
        """Chain in emcee's HDF layout names (``chain`` [nsteps, W, ndim], ``log_prob`` [nsteps, W],
        ``accepted`` [W]) as a compressed ``.npz`` (h5py is not required)."""
        np.savez_compressed(path, chain=self.get_chain(), log_prob=self.get_log_prob(),
                            accepted=self._naccept.cpu().numpy(), iteration=np.array(self._steps_done),
                            **{k: np.asarray(v) for k, v in attrs.items()})


# ----------------------------------------------------------------------------------------------
# iteration schedule of the reference (joxsz_funcs.py:548-635)
# ----------------------------------------------------------------------------------------------

def _generateInitPars(mcmc, fit, rng=None):
    """Initial ball ``p = theta_hat * (1 + N(0, initspread))`` keeping only finite-likelihood draws
    (reference ``joxsz_funcs.py:548-570``); candidates are evaluated a batch at a time."""
    thawedpars = np.array(fit.thawedParVals(), dtype=np.float64)
    assert np.all(np.isfinite(thawedpars))
    walks, dim = mcmc.nwalkers, mcmc.ndim
    rng = np.random if rng is None else rng
    cap = int(getattr(mcmc.engine, "max_walkers", walks))
    p0 = np.empty((0, dim))
    tries = 0
    while p0.shape[0] < walks:
        n = min(cap, max(2 * (walks - p0.shape[0]), 16))
        cand = thawedpars * (1 + rng.normal(0., mcmc.initspread, size=(n, dim)))
        ll = mcmc.engine(cand)
        p0 = np.concatenate([p0, cand[np.isfinite(ll)]], axis=0)
        tries += 1
        if tries > 1000:
            raise RuntimeError("could not draw an initial ensemble with finite likelihood")
    return p0[:walks]


def mcmc_run(mcmc, fit, nburn, nsteps, nthin=1, autorefit=True, minfrac=0.2, minimprove=0.01, max_prefit=None,
             prefit_iterations=1000):
    """MCMC execution with the reference's schedule (``joxsz_funcs.py:572-635``): preliminary 1000-iteration
    rounds repeated while the best log-probability improves, burn-in, then the stored chain.
    ``max_prefit`` optionally bounds the number of preliminary rounds (the reference's loop is unbounded) and
    ``prefit_iterations`` their length (1000 upstream).  With ``chain_shard`` the best log-probability of a round is
    reduced over the ranks, and the restart ensemble is the sampler's full current state."""
    eng = mcmc.engine
    bestprob = float(eng(np.asarray(fit.thawedParVals(), dtype=np.float64)))
    newlike = bestprob
    p0 = _generateInitPars(mcmc, fit)
    print('Preliminary fit (1000 iterations) to improve likelihood')
    rounds = 0
    while newlike >= bestprob:
        bestprob = newlike
        for _ in mcmc.sample(p0, thin=max(prefit_iterations // 2, 1), iterations=prefit_iterations, progress=False):
            pass
        if getattr(mcmc, "chain_shard", False):
            # a sharded chain holds this rank's walkers only; the last stored sample is the current ensemble (an even
            # round length is stored at its last iteration), which every rank holds in full
            newlike = float(mcmc.log_prob_host().max())
            p0 = mcmc.coords_host()
        else:
            newlike = float(mcmc.backend.get_log_prob()[-1, :].max())
            p0 = mcmc.backend.get_chain()[-1, :, :]
        mcmc.backend.reset(mcmc.nwalkers, len(fit.thawedParVals()))
        rounds += 1
        if max_prefit is not None and rounds >= max_prefit:
            break
    print('Burn-in period')
    for _ in mcmc.sample(p0, thin=max(nburn // 2, 1), iterations=nburn, progress=False):
        pass
    p1 = mcmc.backend.get_chain()[-1, :, :] if not getattr(mcmc, "chain_shard", False) else mcmc.coords_host()
    mcmc.backend.reset(mcmc.nwalkers, len(fit.thawedParVals()))
    print('Starting sampling')
    for _ in mcmc.sample(p1, thin=nthin, iterations=nsteps, progress=False):
        pass
    print('Finished sampling')
    print('Acceptance fraction: %s' % np.mean(mcmc.acceptance_fraction))
    return True
