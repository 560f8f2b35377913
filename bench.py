#!/usr/bin/env python
"""bench.py -- log-likelihood evaluations per second of the batched JoXSZ joint SZ + X-ray likelihood.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json, `north_star`): the shipped CL J1226.9+3332 set-up (Nr = 313 radial points,
171 x 171 SZ map, 55 x 55 beam, 19 SZ points, 10 bands x 15 annuli, 13 free parameters) scaled to 65,536
walkers.  One "step" = one ensemble iteration of the stretch-move sampler: every walker's proposal is
evaluated once (two half-ensemble batches, emcee's red/blue split), i.e. 65,536 likelihood evaluations.
With N GPUs the 65,536 evaluations of a step are sharded over the ranks (strong scaling); the only
collective is the all-gather of the accepted half-ensemble after each half-step.

Prints ONE JSON line (rank 0):

* `value`      evaluations / s with the ensemble resident in HBM: the median of `timed_blocks` blocks of EXACTLY K
               steps each (barrier + synchronize on both sides, CUDA events, max over ranks);
* `e2e`        the same metric through the reference-facing call with HOST buffers (N = 1: the vectorised
               `getLikelihood`, pinned host theta in, host log-likelihoods out; N > 1: the sampler iteration with the
               ensemble state copied host -> device before and device -> host after every step, all-gathers inside);
               `e2e_numpy` = `fit.getLikelihood(vals[W, ndim])` with a pageable numpy array, the call emcee makes;
* `roofline`   the dominant kernel (the map stage) against the limit that binds it, the FP64 pipe: executed FP64
               lane-operations / launch time / measured DFMA peak; the HBM form of SURVEY 8(d) is kept under `hbm_form`;
* `cpu_baseline` the oracle's literal per-walker path on the host cores (bounded sample);
* `secondary`  BASELINE configs 3 and 5 (synthetic 255- and 511-pixel clusters), 8,192 walkers per rank, outside the
               headline timed region;
* `state_checksum` 64-bit sums of the ensemble after the timed blocks: equal for every N (rank-count invariance).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "log-likelihood evals/sec (walker-steps/s)"
UNIT = "evals/s"
TOTAL_WALKERS = 65536
TIMED_BLOCKS = 5


# ----------------------------------------------------------------------------------------------
# shared set-up
# ----------------------------------------------------------------------------------------------

WORKLOAD = "cl1226"          # set from --workload before anything is built (also in the CPU pool workers)

WORKLOADS = {
    # name: (map_half, nr) of joxsz_b200.cluster.synthetic_inputs; None = the shipped cluster
    "cl1226": None,
    "synth255": (127, 512),      # BASELINE config 3: 512-point grid, 255-pixel map (nearest odd side to 256)
    "synth511": (255, 1024),     # BASELINE config 5: 1024-point grid, 511-pixel map (nearest odd side to 512)
}


def build_cluster(workload=None):
    from joxsz_b200 import cluster
    from joxsz_b200.mb import mb
    mb.fit.debugfit = False
    workload = workload or WORKLOAD
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    if WORKLOADS[workload] is not None:
        map_half, nr = WORKLOADS[workload]
        inp = cluster.synthetic_inputs(map_half=map_half, nr=nr, base=inp)
    fit, _ = cluster.build_fit(inp, savedir=None)
    return fit


def workload_config(walkers, n_gpus, workload=None):
    """The `config` object: identical for the GPU arm and the reference arm of the same command line."""
    workload = workload or WORKLOAD
    if workload == "cl1226":
        cfg = {"workload": "CL J1226.9+3332 (shipped example) scaled to 65,536 walkers: Nr=313, map 171x171, "
                           "beam 55x55, 19 SZ points, 10 bands x 15 annuli, 13 free parameters; "
                           "step = one stretch-move ensemble iteration (65,536 likelihood evaluations)",
               "walkers": walkers, "nr": 313, "map": 171, "ndim": 13}
    else:
        map_half, nr = WORKLOADS[workload]
        n = 2 * map_half + 1
        cfg = {"workload": f"synthetic cluster (joxsz_b200.cluster.synthetic_inputs): Nr={nr}, map {n}x{n} (the reference "
                           f"builds odd sides only), Gaussian beam 55x55, normal-cdf transfer function, shipped X-ray "
                           f"layout; step = one stretch-move ensemble iteration",
               "walkers": walkers, "nr": nr, "map": n, "ndim": 13}
    cfg.update({"xray_tables": "synthetic (XSPEC unavailable)",
                "l2_policy": "inputs and intermediates per step (> 400 MB) exceed the 126 MB L2; no flush needed",
                "parallelism": f"walkers sharded over {n_gpus} GPU(s)"})
    return cfg


def ensemble(fit, n, seed=20260103):
    """Valid (finite-likelihood) starting ensemble: a tight ball around the fiducial parameters."""
    from joxsz_b200.synthetic import draw_parameters
    return draw_parameters(fit.thawed, n=n, seed=seed, spread=0.02, frac_bad=0.0)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._proc = None

    def __enter__(self):
        # one long-lived nvidia-smi sampling every 50 ms (spawning it per sample is too slow for short runs)
        try:
            self._proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                           "--format=csv,noheader,nounits", "-lms", "50"],
                                          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)           # first sample lands before the timed region starts
        except Exception:
            self._proc = None
        return self

    def __exit__(self, *a):
        if self._proc is None:
            return
        time.sleep(0.06)
        self._proc.terminate()
        try:
            out, _ = self._proc.communicate(timeout=5)
        except Exception:
            self._proc.kill()
            out = ""
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) >= 7:
                self.samples.append(f)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's per-walker path)
# ----------------------------------------------------------------------------------------------

_ORACLE_SETUPS = {}


def _oracle_setup(workload):
    s = _ORACLE_SETUPS.get(workload)
    if s is None:
        from helpers import oracle_setup_from_fit
        s = _ORACLE_SETUPS[workload] = oracle_setup_from_fit(build_cluster(workload))
    return s


def _cpu_init():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def _cpu_prepare(workload):
    _oracle_setup(workload)
    time.sleep(0.25)          # long enough that every worker of the pool takes exactly one of these tasks
    return os.getpid()


def _cpu_eval(task):
    from oracle import joxsz_oracle as orc
    workload, theta = task
    return orc.get_likelihood(theta, _oracle_setup(workload))


def cpu_reference_rate(thetas, pool, workload=None):
    """evals/s of the literal per-walker path, one task per walker like emcee's pool.map (joxsz_main.py:203-208)."""
    workload = workload or WORKLOAD
    pool.map(_cpu_prepare, [workload] * pool._processes, chunksize=1)     # set-ups built before the clock starts
    t0 = time.perf_counter()
    out = pool.map(_cpu_eval, [(workload, th) for th in thetas], chunksize=1)
    dt = time.perf_counter() - t0
    return len(thetas) / dt, dt, np.array(out)


def make_pool():
    """Process pool over all host cores.  Created BEFORE the process touches CUDA / NCCL: forking a multi-threaded
    parent later can leave a child stuck on a lock that some other thread held at fork time."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores, initializer=_cpu_init)
    pool.map(_noop, range(cores * 2))
    return pool, cores


def _noop(x):
    return x


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool, cores = make_pool()
    fit = build_cluster()
    per_step = max(cores * 16, 128)
    thetas = ensemble(fit, per_step * (args.steps + args.warmup))
    for w in range(args.warmup):
        cpu_reference_rate(thetas[w * per_step:(w + 1) * per_step], pool)
    t_tot = 0.0
    for k in range(args.steps):
        lo = (args.warmup + k) * per_step
        _, dt, _ = cpu_reference_rate(thetas[lo:lo + per_step], pool)
        t_tot += dt
    pool.close()
    rate = per_step * args.steps / t_tot
    sample = f"{per_step} walkers per step of the {args.walkers}-walker workload, literal per-walker path, Pool({cores})"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.walkers, args.gpus),
            "reference_arm": "oracle port of joxsz_funcs.getLikelihood on the host cores (the Python reference and its "
                             "dependencies cannot travel to the GPU box); each step evaluates a bounded sample of the "
                             "workload's walkers",
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def measured_fp64_peaks(device):
    """FP64 FMA / DMMA peaks of this GPU from the microbenchmarks in scripts/libjx_peaks.so (measurement tooling,
    not part of the product library)."""
    from joxsz_b200 import build as jb
    lib = C.CDLL(jb.build_peaks())
    out = {}
    for key, fn in (("dfma_tflops", "jxp_measure_fp64_tflops"), ("dmma_tflops", "jxp_measure_dmma_tflops")):
        v = C.c_double(0.0)
        f = getattr(lib, fn)
        f.argtypes = [C.c_int32, C.POINTER(C.c_double)]
        f.restype = C.c_int
        out[key] = v.value if f(device, C.byref(v)) != 0 else v.value
    return out


def fp64_lane_ops(pk):
    """Executed FP64 lane-operations (one DFMA / DADD / DMUL of one thread, counting the lanes a warp instruction
    occupies) of the map kernel per walker, from the kernel's structure (cross-checked against
    ncu's smsp__inst_executed_pipe_fp64 in profiles/)."""
    H, P, nbeam = pk.H, pk.map_ops.P, int(pk.bmix.shape[0])
    Q = P // 2 + 1
    npair = (H + 1) // 2
    fft9 = 9 * 2 * 230          # nine threads x (two DFT-16 + twiddles) ~ 230 FP64 instructions each per pass pair
    if P == 256:
        synth = (H * (H + 1) // 2) * 4
        rows = 2 * npair * fft9 * 32 // 27               # 27 of 32 lanes carry a transform
        ycols = 512 * 22 * (2 * nbeam - 1)               # 512 threads x 22 rows x (2 nbeam - 1) taps
        return synth + rows + ycols
    R = P // 256
    synth = (H * (H + 1) // 2) * 4
    ns = 3 if R == 4 else R                              # branches computed (the s = 3 branch of radix 4 is mirrored)
    rows = 2 * npair * ns * 16 * 2 * 230                 # sixteen-thread full-complex FFT-256 x ns branches per ROW PAIR
    ycols = ((H + 15) // 16 * 16) * Q * (2 * nbeam - 1)
    return synth + rows + ycols


def large_map_kernel_name(pk):
    """Which kernel jx_create picks for cyclic lengths 512 / 1024 (jx_szmap_large2_ok in csrc/k3l2_szmap.cu)."""
    k3l2 = os.environ.get("JX_K3L2", "1") != "0" and int(pk.bmix.shape[0]) <= 28
    return "k3l2_szmap_kernel" if k3l2 else "k3l_szmap_kernel"


def note(msg):
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def time_blocks(step_fn, steps, blocks, barrier, reduce_max):
    import torch
    out = []
    for _ in range(blocks):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        out.append(reduce_max(e0.elapsed_time(e1)))
    return out


def state_checksum(sampler):
    """64-bit wrapping sums of the bit patterns of the ensemble (positions, log-probs) and of the accept counters."""
    import torch
    c = sampler._coords.view(torch.int64).sum().item() & 0xFFFFFFFFFFFFFFFF
    l = sampler._lp.view(torch.int64).sum().item() & 0xFFFFFFFFFFFFFFFF
    a = int(sampler._naccept.sum().item())
    return {"coords": f"{c:016x}", "log_prob": f"{l:016x}", "accepted": a, "iteration": sampler.iteration}


def profile_pass(eng, sampler, steps):
    """Per-kernel CUDA-event timers need kernel-by-kernel launches: `steps` more iterations of the same chain with the
    graph suspended, right after the timed blocks.  Returns {stage: (total_ms, launches)}."""
    import torch
    eng.set_profiling(True)
    eng.stage_times()
    sampler.eager_only = True
    for _ in range(steps):
        sampler.step()
    torch.cuda.synchronize()
    st = eng.stage_times()
    eng.set_profiling(False)
    sampler.eager_only = False
    return st


def run_secondary(name, world, rank, local, dist, args, barrier, reduce_max, pool, cores):
    """BASELINE configs 3 / 5: 8,192 walkers per rank of a larger synthetic cluster, a few iterations."""
    import torch
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.sampler import EnsembleSampler
    t_build = time.perf_counter()
    note(f"secondary {name}: building the cluster")
    fit = build_cluster(name)
    Wn = args.secondary_walkers * world
    eng = BatchedLikelihood(fit, max_walkers=Wn // world + 64, device=local)
    sampler = EnsembleSampler(Wn, eng.ndim, eng, seed=4321, world_size=world, rank=rank,
                              group=(dist.group.WORLD if world > 1 else None))
    note(f"secondary {name}: engine ready, initialising {Wn} walkers")
    sampler.initialize(ensemble(fit, Wn, seed=20260105))
    t_build = time.perf_counter() - t_build
    note(f"secondary {name}: set-up {t_build:.1f} s, stepping")
    steps = args.secondary_steps
    for _ in range(3):
        sampler.step()
    blocks = time_blocks(sampler.step, steps, 3, barrier, reduce_max)
    ms = sorted(blocks)[len(blocks) // 2]
    chk = state_checksum(sampler)
    note(f"secondary {name}: {ms / steps:.2f} ms / step, profiling pass")
    stages = profile_pass(eng, sampler, 2)
    note(f"secondary {name}: CPU sample")
    out = None
    if rank == 0:
        pk = eng.packed
        ncalls = max(stages["szmap"][1], 1)
        stage_ms = {k: v[0] / ncalls for k, v in stages.items()}
        nw = sampler.evals_per_rank_per_launch()
        lane = fp64_lane_ops(pk)
        k3_s = stage_ms["szmap"] * 1e-3
        # parity of the device path against the literal per-walker oracle on a small sample of the final ensemble
        theta = sampler.coords_host()[:args.secondary_cpu_sample]
        cpu_rate, cpu_dt, cpu_ll = cpu_reference_rate(theta, pool, name)
        gpu_ll = eng(theta)
        fin = np.isfinite(cpu_ll)
        out = {"config": workload_config(Wn, world, name), "value": Wn * steps / (ms * 1e-3), "unit": UNIT,
               "ms_per_step": ms / steps, "steps": steps, "timed_blocks": len(blocks),
               "block_ms": [b / steps for b in blocks], "graph": sampler.graph_active,
               "map_kernel": large_map_kernel_name(pk) if pk.map_ops.P != 256 else "k3w_szmap_kernel",
               "stage_ms_per_launch": stage_ms, "walkers_per_launch": nw,
               "map_kernel_fp64_lane_ops_per_walker": lane,
               "map_kernel_fp64_lane_ops_per_s": lane * nw / k3_s if k3_s > 0 else None,
               "map_kernel_hbm_form_gbs": pk.algorithmic_bytes()["szmap"] * nw / k3_s / 1e9 if k3_s > 0 else None,
               "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{len(theta)} walkers, literal per-walker oracle path, {cpu_dt:.1f} s"},
               "parity_max_abs_dll_vs_cpu_sample": float(np.max(np.abs(gpu_ll[fin] - cpu_ll[fin]))) if fin.any() else None,
               "parity_inf_mask_equal": bool(np.array_equal(fin, np.isfinite(gpu_ll))),
               "state_checksum": chk, "setup_s": t_build,
               "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9}
    sampler.close()
    eng.close()
    del sampler, eng
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    pool, cores = make_pool() if rank == 0 else (None, 0)      # before CUDA is initialised (see make_pool)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from joxsz_b200 import funcs
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.sampler import EnsembleSampler

    note("building the cluster")
    fit = build_cluster()
    W = args.walkers
    eng = BatchedLikelihood(fit, max_walkers=max(W // world + 64, 1024) if world > 1 else W, device=local)
    note("engine ready")
    p0 = ensemble(fit, W)
    sampler = EnsembleSampler(W, eng.ndim, eng, seed=1234, world_size=world, rank=rank,
                              group=(dist.group.WORLD if world > 1 else None), graph=not args.no_graph,
                              exchange=args.exchange)
    sampler.initialize(p0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    note("sampler initialised, warm-up")
    for _ in range(args.warmup):
        sampler.step()
    extra_warm = 0
    while sampler.use_graph and not sampler.graph_active and sampler._graph_failed is None and extra_warm < 4:
        sampler.step()                       # the iteration is captured after its first eager steps
        extra_warm += 1
    note(f"timed blocks (graph {sampler.graph_active}, p2p {sampler._px is not None})")
    with ClockSampler(local) as clk:
        blocks = time_blocks(sampler.step, args.steps, TIMED_BLOCKS, barrier, reduce_max)
    ms = sorted(blocks)[len(blocks) // 2]
    value = W * args.steps / (ms * 1e-3)
    acc = sampler.mean_acceptance()
    checksum = state_checksum(sampler)
    note(f"{ms / args.steps:.3f} ms / step; profiling pass")
    stages = profile_pass(eng, sampler, args.steps)
    note("end-to-end legs")

    # ---- end to end with host buffers
    shard = W // world
    host_state = sampler.coords_host()
    if world == 1:
        host_theta = host_state.copy()
        # the step's inputs live in pinned host memory (the caller's buffer); every call copies them to the device,
        # runs the kernels and reads the log-likelihoods back
        host_theta_pinned = torch.from_numpy(host_theta).pin_memory()
        out_holder = {}

        def e2e_step():
            out_holder["ll"] = eng(host_theta_pinned).numpy()

        for _ in range(2):
            e2e_step()
        t0 = time.perf_counter()
        e2e_blocks = time_blocks(e2e_step, args.steps, 3, barrier, reduce_max)
        e2e_ms = sorted(e2e_blocks)[1]
        assert np.isfinite(out_holder["ll"]).any()
        e2e = {"value": W * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(W * eng.ndim * 8),
               "d2h_bytes_per_step": int(W * 8), "block_ms": [b / args.steps for b in e2e_blocks],
               "call": "BatchedLikelihood.__call__(pinned host theta[65536, 13]) -> host ll (the vectorised getLikelihood)"}
        # the call emcee makes (joxsz_main.py:206): the bound fit.getLikelihood on a pageable numpy array
        funcs.attach_engine(fit, eng)
        vals = host_theta.copy()

        def e2e_numpy_step():
            out_holder["ll2"] = fit.getLikelihood(vals)

        for _ in range(2):
            e2e_numpy_step()
        nb = time_blocks(e2e_numpy_step, args.steps, 3, barrier, reduce_max)
        assert np.array_equal(out_holder["ll2"], out_holder["ll"])
        funcs.detach_engine(fit)
        e2e_numpy = {"value": W * args.steps / (sorted(nb)[1] * 1e-3), "unit": UNIT,
                     "call": "fit.getLikelihood(vals[65536, 13]) with a pageable numpy array (staged through the engine's "
                             "pinned buffer), host array out"}
        note("collapsed mode")
        # optional `ll`-only mode (never the headline): the linear SZ chain as one operator, same engine
        theta_dev = torch.from_numpy(host_theta).to(eng.device)
        ll_staged = eng.loglike_device(theta_dev).clone()
        eng.mode = "collapsed"
        ll_coll = eng.loglike_device(theta_dev).clone()

        def coll_step():
            eng.loglike_device(theta_dev, out=ll_coll)

        cb = time_blocks(coll_step, args.steps, 3, barrier, reduce_max)
        eng.mode = "staged"
        finc = torch.isfinite(ll_staged)
        collapsed = {"value": W * args.steps / (sorted(cb)[1] * 1e-3), "unit": UNIT, "ms_per_call": sorted(cb)[1] / args.steps,
                     "what": "jx_loglike_collapsed on the device-resident ensemble (K1 -> one DMMA GEMM with the operator "
                             "built from the staged kernels -> K5); intermediate maps do not exist in this mode",
                     "max_abs_dll_vs_staged": float((ll_staged[finc] - ll_coll[finc]).abs().max().item()),
                     "inf_mask_equal": bool(torch.equal(finc, torch.isfinite(ll_coll)))}
    else:
        collapsed = None
        # N > 1: the sampler iteration with the ensemble state in host memory between steps -- H2D of positions and
        # log-probs before, the two half-steps with their all-gathers, D2H of the new state after
        pin_c = torch.from_numpy(host_state.copy()).pin_memory()
        pin_l = torch.from_numpy(sampler.log_prob_host().copy()).pin_memory()

        def e2e_step():
            sampler._coords.copy_(pin_c, non_blocking=True)
            sampler._lp.copy_(pin_l, non_blocking=True)
            sampler.step()
            pin_c.copy_(sampler._coords, non_blocking=True)
            pin_l.copy_(sampler._lp, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            e2e_step()
        e2e_blocks = time_blocks(e2e_step, args.steps, 3, barrier, reduce_max)
        e2e_ms = sorted(e2e_blocks)[1]
        nbytes = int(W * (eng.ndim + 1) * 8)
        e2e = {"value": W * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "block_ms": [b / args.steps for b in e2e_blocks],
               "call": "EnsembleSampler.step() with the ensemble state (positions, log-probs) copied from pinned host "
                       "memory before and back after every iteration, on every rank; all-gathers inside"}
        e2e_numpy = None
        host_theta = host_state[rank * shard:(rank + 1) * shard].copy()

    note("roofline / CPU baseline")
    line = None
    if rank == 0:
        pk = eng.packed
        alg = pk.algorithmic_bytes()
        flops = pk.algorithmic_flops()
        k3_ms, k3_n = stages["szmap"]
        nw = sampler.evals_per_rank_per_launch()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fp = measured_fp64_peaks(local)
        k3_s = (k3_ms / max(k3_n, 1)) * 1e-3
        lane_model = fp64_lane_ops(pk)
        lane, lane_src, tr = lane_model, "counted from the kernel's structure (bench.fp64_lane_ops)", None
        map_kernel = "k3_szmap_kernel" if pk.map_ops.P == 256 else large_map_kernel_name(pk)
        try:
            if WORKLOAD == "cl1226":
                tr = json.load(open(os.path.join(ROOT, "profiles", "k3_ncu_traffic.json")))
                map_kernel = tr.get("kernel", map_kernel)
                if "fp64_lane_ops_per_walker" in tr:           # the executed count, from ncu's FP64-pipe instruction counter
                    lane, lane_src = tr["fp64_lane_ops_per_walker"], "ncu: " + tr.get("fp64_lane_ops_how", "")
        except Exception:
            tr = None
        lane_rate = lane * nw / k3_s if k3_n else None
        lane_peak = fp["dfma_tflops"] * 1e12 / 2.0            # one DFMA = 2 flop = one lane-operation
        hbm_gbs = alg["szmap"] * nw / k3_s / 1e9 if k3_n else None
        roof = {"bound": "fp64", "kernel": map_kernel,
                "achieved": 2.0 * lane_rate / 1e12 if lane_rate else None, "peak": fp["dfma_tflops"], "unit": "TFLOP/s",
                "frac": lane_rate / lane_peak if lane_rate and lane_peak else None, "traffic": None,
                "peak_source": "DFMA microbenchmark in this run (scripts/jx_peaks.cu); MEASURED_PEAKS.json has no FP64 figure",
                "executed_fp64_lane_ops_per_walker": lane, "lane_ops_source": lane_src,
                "lane_ops_per_walker_structural_count": lane_model, "walkers_per_launch": nw,
                "practical_ceiling": "a DFMA stream with two fresh 64-bit sources per instruction (the y convolution's "
                                     "pattern) issues at 0.38-0.41 of a warp-instruction / clock / scheduler on B200 against "
                                     "0.487 for the peak pattern a=fma(a,m,c): scripts/dfma_pattern_microbench.cu, "
                                     "profiles/r02_results.md",
                "avg_launch_ms": k3_s * 1e3, "launches_timed": int(k3_n),
                "timing": "CUDA events around every kernel in an eager pass of `steps` iterations right after the timed "
                          "blocks (the timed blocks replay one CUDA graph per iteration, which has no per-kernel events)",
                "note": "achieved = executed FP64 lane-operations x 2 flop / launch time; the kernel keeps the maps in "
                        "shared memory, so its DRAM traffic is a few per cent of the staged-reference bytes and HBM is "
                        "not what binds it (hbm_form, kept for SURVEY 8d's convention)",
                "hbm_form": {"alg_bytes_per_walker": alg["szmap"], "achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak,
                             "frac": hbm_gbs / hbm_peak if hbm_gbs else None, "peak_source": peak_src},
                "alg_flops_per_walker_survey_8d": flops["szmap"]}
        if tr is not None:
            roof["traffic"] = tr["dram_bytes_per_launch"] * (nw / tr["walkers_per_launch"])
            roof["traffic_source"] = tr.get("source")
            roof["ncu"] = {k: tr[k] for k in ("fp64_pipe_busy_frac", "shared_pipe_busy_frac", "issue_slots_busy_frac",
                                              "fp64_lane_ops_per_walker", "gpu_time_ms_under_ncu") if k in tr}
        ncalls = max(k3_n, 1)
        stage_ms = {k: (v[0] / ncalls) for k, v in stages.items()}

        def hbm(stage, key):
            t_s = stage_ms[stage] * 1e-3
            gbs = alg[key] * nw / t_s / 1e9 if t_s > 0 else None
            return {"bound": "hbm", "alg_bytes_per_walker": alg[key], "ms": stage_ms[stage], "achieved_gbs": gbs,
                    "frac_of_measured_hbm": gbs / hbm_peak if gbs else None}

        def tensor(stage, key):
            t_s = stage_ms.get(stage, 0.0) * 1e-3
            tf = flops[key] * nw / t_s / 1e12 if t_s > 0 else None
            return {"bound": "tensor(fp64 dmma)", "alg_flops_per_walker": flops[key], "ms": stage_ms.get(stage),
                    "achieved_tflops": tf, "peak_tflops_measured_dmma": fp["dmma_tflops"],
                    "frac": tf / fp["dmma_tflops"] if tf and fp["dmma_tflops"] else None}

        both_s = (stage_ms["szmap"] + stage_ms.get("filter", 0.0)) * 1e-3
        both_gbs = alg["szmap"] * nw / both_s / 1e9 if both_s > 0 else None
        stage_roof = {"profiles": hbm("profiles", "profiles"),
                      "szmap": {"bound": "fp64", "ms": stage_ms["szmap"], "frac": roof["frac"]},
                      "xray": dict(hbm("xray", "xray"), note="side stream: elapsed time overlaps the project / szmap stages"),
                      "tail": hbm("tail", "tail"), "filter": tensor("filter", "filter"), "project": tensor("project", "project"),
                      "szmap_plus_filter": {"bound": "hbm", "alg_bytes_per_walker": alg["szmap"], "ms": both_s * 1e3,
                                            "achieved_gbs": both_gbs,
                                            "frac_of_measured_hbm": both_gbs / hbm_peak if both_gbs else None}}
        # bounded CPU baseline on this box's cores + parity of the device path on the same walkers
        sample_n = max(cores * 64, 512)
        cpu_rate, cpu_dt, cpu_ll = cpu_reference_rate(host_theta[:sample_n], pool)
        gpu_ll = eng(host_theta[:sample_n])
        fin = np.isfinite(cpu_ll)
        parity = float(np.max(np.abs(gpu_ll[fin] - cpu_ll[fin]))) if fin.any() else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(W, world),
                "timed_blocks": TIMED_BLOCKS, "block_ms_per_step": [b / args.steps for b in blocks],
                "value_is": "median block", "warmup_extra_steps": extra_warm,
                "acceptance_fraction": acc, "state_checksum": checksum,
                "sampler": {"cuda_graph": sampler.graph_active, "graph_fallback_reason": sampler._graph_failed,
                            "exchange": ("none (1 rank)" if world == 1 else
                                         "p2p: accept kernel stores into every rank's buffer over NVLink, flags, no collective"
                                         if sampler._px is not None else "nccl all_gather_into_tensor per half-step"),
                            "p2p_fallback_reason": sampler._px_failed},
                "clocks": clk.summary(), "e2e": e2e, "e2e_numpy": e2e_numpy, "collapsed_mode": collapsed,
                "gpu_launches": int(sampler.launches_per_step() * args.steps),
                "gpu_launches_note": "kernels of libjoxsz_b200.so per timed block (K steps), replayed from one CUDA graph "
                                     "per iteration: 2 x (K1 K4 K2 K3 K7 K5 + propose accept scatter) + permutation",
                "roofline": roof, "stage_ms_per_launch": stage_ms, "stage_rooflines": stage_roof,
                "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{sample_n} walkers of the same ensemble, literal per-walker oracle path, "
                                           f"Pool({cores}), {cpu_dt:.1f} s"},
                "parity_max_abs_dll_vs_cpu_sample": parity}
    sampler.close()             # the captured iteration (with its collectives) goes before the process group does
    eng.close()
    del sampler, eng
    torch.cuda.empty_cache()

    if args.secondary and WORKLOAD == "cl1226":
        sec = {}
        for name in ("synth255", "synth511"):
            try:
                r = run_secondary(name, world, rank, local, dist, args, barrier, reduce_max, pool, cores)
            except Exception as e:          # the headline line must still be printed
                r = {"error": f"{type(e).__name__}: {e}"}
            if rank == 0:
                sec[name] = r
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        pool.close()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=TOTAL_WALKERS)
    ap.add_argument("--workload", default="cl1226", choices=sorted(WORKLOADS),
                    help="cl1226 = the configuration the metric is quoted on (default); synth255 / synth511 = the "
                         "larger synthetic clusters of BASELINE configs 3 and 5")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the BASELINE config 3 / 5 block")
    ap.add_argument("--secondary-walkers", type=int, default=8192, help="walkers per rank in the secondary block")
    ap.add_argument("--secondary-steps", type=int, default=3)
    ap.add_argument("--secondary-cpu-sample", type=int, default=32)
    ap.add_argument("--no-graph", action="store_true", help="launch the sampler iteration kernel by kernel")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N > 1: how the half-step results travel between the ranks (see EnsembleSampler)")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
