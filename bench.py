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

Prints ONE JSON line (rank 0).  `value` = evaluations / s with the ensemble resident in HBM;
`e2e` = the same metric through the reference-facing call (`BatchedLikelihood.__call__`, the vectorised
`getLikelihood`) with host numpy buffers in and out; `roofline` describes the dominant kernel (the map
stage K3); `cpu_baseline` is the oracle's literal per-walker path on the host cores.
"""
from __future__ import annotations

import argparse
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


def build_cluster():
    from joxsz_b200 import cluster
    from joxsz_b200.mb import mb
    mb.fit.debugfit = False
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    if WORKLOADS[WORKLOAD] is not None:
        map_half, nr = WORKLOADS[WORKLOAD]
        inp = cluster.synthetic_inputs(map_half=map_half, nr=nr, base=inp)
    fit, _ = cluster.build_fit(inp, savedir=None)
    return fit


def workload_config(extra=None):
    if WORKLOAD == "cl1226":
        cfg = {"workload": "CL J1226.9+3332 (shipped example) scaled to 65,536 walkers: Nr=313, map 171x171, "
                           "beam 55x55, 19 SZ points, 10 bands x 15 annuli, 13 free parameters; "
                           "step = one stretch-move ensemble iteration (65,536 likelihood evaluations)",
               "walkers": TOTAL_WALKERS, "nr": 313, "map": 171, "ndim": 13}
    else:
        map_half, nr = WORKLOADS[WORKLOAD]
        n = 2 * map_half + 1
        cfg = {"workload": f"synthetic cluster (joxsz_b200.cluster.synthetic_inputs): Nr={nr}, map {n}x{n} (the reference "
                           f"builds odd sides only), Gaussian beam 55x55, normal-cdf transfer function, shipped X-ray "
                           f"layout; step = one stretch-move ensemble iteration",
               "walkers": TOTAL_WALKERS, "nr": nr, "map": n, "ndim": 13}
    cfg.update({"xray_tables": "synthetic (XSPEC unavailable)",
                "l2_policy": "inputs and intermediates per step (> 400 MB) exceed the 126 MB L2; no flush needed"})
    if extra:
        cfg.update(extra)
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

_ORACLE_SETUP = None


def _cpu_init(workload="cl1226"):
    global _ORACLE_SETUP, WORKLOAD
    WORKLOAD = workload
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from helpers import oracle_setup_from_fit
    _ORACLE_SETUP = oracle_setup_from_fit(build_cluster())


def _cpu_eval(theta):
    from oracle import joxsz_oracle as orc
    return orc.get_likelihood(theta, _ORACLE_SETUP)


def cpu_reference_rate(thetas, pool):
    """evals/s of the literal per-walker path, one task per walker like emcee's pool.map (joxsz_main.py:203-208)."""
    t0 = time.perf_counter()
    out = pool.map(_cpu_eval, list(thetas), chunksize=1)
    dt = time.perf_counter() - t0
    return len(thetas) / dt, dt, np.array(out)


def make_pool():
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores, initializer=_cpu_init, initargs=(WORKLOAD,))
    pool.map(_noop, range(cores * 2))
    return pool, cores


def _noop(x):
    return x


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fit = build_cluster()
    pool, cores = make_pool()
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
    sample = f"{per_step} walkers per step of the 65,536-walker workload, literal per-walker path, Pool({cores})"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config({"reference_arm": "oracle port of joxsz_funcs.getLikelihood on host cores "
                                                        "(the Python reference and its dependencies cannot travel)"}),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.sampler import EnsembleSampler

    fit = build_cluster()
    W = args.walkers
    eng = BatchedLikelihood(fit, max_walkers=max(W // world + 64, 1024), device=local)
    p0 = ensemble(fit, W)
    sampler = EnsembleSampler(W, eng.ndim, eng, seed=1234, world_size=world, rank=rank,
                              group=(dist.group.WORLD if world > 1 else None))
    sampler.initialize(p0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sampler.step()
    eng.set_profiling(True)
    eng.stage_times()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            sampler.step()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    stages = eng.stage_times()
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = W * args.steps / (ms * 1e-3)
    acc = sampler.mean_acceptance()

    # ---- end to end through the reference-facing vectorised call with host buffers (rank-local shard)
    shard = W // world
    host_theta = sampler.coords_host()[rank * shard:(rank + 1) * shard].copy()
    # the step's inputs live in pinned host memory (the caller's buffer); every call copies them to the device,
    # runs the kernels and reads the log-likelihoods back
    host_theta_pinned = torch.from_numpy(host_theta).pin_memory()
    for _ in range(2):
        eng(host_theta_pinned)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out = eng(host_theta_pinned).numpy()     # pinned host tensor in (H2D), kernels, D2H, host array out -- every step
    e1.record()
    torch.cuda.synchronize()
    e2e_wall_s = time.perf_counter() - t0
    e2e_s = max(e0.elapsed_time(e1) * 1e-3, e2e_wall_s)      # device clock; the host clock can only be longer
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = shard * world * args.steps / float(t.item())
    assert np.isfinite(out).any()

    if rank == 0:
        pk = eng.packed
        alg = pk.algorithmic_bytes()
        k3_ms, k3_n = stages["szmap"]
        k3_walkers = sampler.evals_per_rank_per_launch()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        k3_avg_s = (k3_ms / max(k3_n, 1)) * 1e-3
        achieved = alg["szmap"] * k3_walkers / k3_avg_s / 1e9 if k3_n else None
        import ctypes as C
        tf = C.c_double(0.0)
        eng.lib.jx_measure_fp64_tflops(local, C.byref(tf))
        tfd = C.c_double(0.0)
        eng.lib.jx_measure_dmma_tflops(local, C.byref(tfd))
        flops = pk.algorithmic_flops()
        roof = {"bound": "hbm", "kernel": "k3_szmap_kernel" if WORKLOAD == "cl1226" else "k3l_szmap_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": (achieved / hbm_peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_walker": alg["szmap"], "walkers_per_launch": k3_walkers,
                "avg_launch_ms": k3_avg_s * 1e3, "launches_timed": int(k3_n),
                "note": "algorithmic bytes = the maps of the stages this kernel replaces, materialised once (write y_2d, "
                        "read y_2d, write the convolved map; SURVEY 8d convention); the kernel keeps them in shared "
                        "memory and writes only the distinct pixels of the convolved map, so DRAM traffic is far below "
                        "this and the binding limit is FP64 throughput.  The filter stage is the K7 GEMM "
                        "(stage_rooflines.filter); stage_rooflines.szmap_plus_filter has both against the same bytes",
                "fp64": {"alg_flops_per_walker": flops["szmap"], "measured_dfma_peak_tflops": tf.value, "measured_dmma_peak_tflops": tfd.value,
                         "achieved_tflops": flops["szmap"] * k3_walkers / k3_avg_s / 1e12 if k3_n else None}}
        ncalls = max(k3_n, 1)
        stage_ms = {k: (v[0] / ncalls) for k, v in stages.items()}
        # DRAM traffic of the dominant kernel from the committed ncu capture of this same command (per launch)
        try:
            if WORKLOAD != "cl1226":
                raise KeyError("no ncu capture for this workload")
            tr = json.load(open(os.path.join(ROOT, "profiles", "k3_ncu_traffic.json")))
            roof["traffic"] = tr["dram_bytes_per_launch"] * (k3_walkers / tr["walkers_per_launch"])
            roof["traffic_source"] = tr.get("source")
            # what the kernel is actually bound by, from the same ncu capture (pipe-busy fractions of the SMs)
            roof["executed"] = {k: tr[k] for k in ("fp64_pipe_busy_frac", "shared_pipe_busy_frac",
                                                   "issue_slots_busy_frac") if k in tr}
        except Exception:
            pass
        # every stage against the roofline north_star names for it: HBM for profiles / map / reduction,
        # FP64 tensor (DMMA) peak for the projection GEMM.  Algorithmic figures: SURVEY.md 8(d), DESIGN.md 5.
        nw = k3_walkers
        def hbm(stage, key):
            t_s = stage_ms[stage] * 1e-3
            gbs = alg[key] * nw / t_s / 1e9 if t_s > 0 else None
            return {"bound": "hbm", "alg_bytes_per_walker": alg[key], "ms": stage_ms[stage], "achieved_gbs": gbs,
                    "frac_of_measured_hbm": gbs / hbm_peak if gbs else None}
        proj_tf = flops["project"] * nw / (stage_ms["project"] * 1e-3) / 1e12 if stage_ms["project"] > 0 else None
        filt_tf = flops["filter"] * nw / (stage_ms["filter"] * 1e-3) / 1e12 if stage_ms.get("filter", 0) > 0 else None
        both_s = (stage_ms["szmap"] + stage_ms.get("filter", 0.0)) * 1e-3
        both_gbs = alg["szmap"] * nw / both_s / 1e9 if both_s > 0 else None
        stage_roof = {"profiles": hbm("profiles", "profiles"), "szmap": hbm("szmap", "szmap"),
                      "xray": dict(hbm("xray", "xray"), note="side stream: elapsed time overlaps the project / szmap stages "
                                   "(0.105 ms per 32768 walkers when run alone)"),
                      "tail": hbm("tail", "tail"),
                      # the transfer-function filter is a GEMM over the walkers (K7) when the cyclic length is 256
                      "filter": {"bound": "tensor(fp64 dmma)", "alg_flops_per_walker": flops["filter"],
                                 "ms": stage_ms.get("filter"), "achieved_tflops": filt_tf,
                                 "peak_tflops_measured_dmma": tfd.value,
                                 "frac": filt_tf / tfd.value if filt_tf and tfd.value else None},
                      # north_star item (3) as a whole: map synthesis + beam convolution + filtering
                      "szmap_plus_filter": {"bound": "hbm", "alg_bytes_per_walker": alg["szmap"], "ms": both_s * 1e3,
                                            "achieved_gbs": both_gbs,
                                            "frac_of_measured_hbm": both_gbs / hbm_peak if both_gbs else None},
                      "project": {"bound": "tensor(fp64 dmma)", "alg_flops_per_walker": flops["project"],
                                  "ms": stage_ms["project"], "achieved_tflops": proj_tf,
                                  "peak_tflops_measured_dmma": tfd.value,
                                  "frac": proj_tf / tfd.value if proj_tf and tfd.value else None}}
        launches = int(sum(v[1] for v in stages.values())) + sampler.aux_launches
        # bounded CPU baseline on this box's cores
        pool, cores = make_pool()
        sample_n = max(cores * 64, 512)
        cpu_rate, cpu_dt, cpu_ll = cpu_reference_rate(host_theta[:sample_n], pool)
        pool.close()
        gpu_ll = eng(host_theta[:sample_n])
        fin = np.isfinite(cpu_ll)
        parity = float(np.max(np.abs(gpu_ll[fin] - cpu_ll[fin]))) if fin.any() else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config({"walkers": W, "parallelism": f"walkers sharded over {world} GPU(s)",
                                           "acceptance_fraction": acc}),
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(shard * eng.ndim * 8),
                        "d2h_bytes_per_step": int(shard * 8),
                        "call": "BatchedLikelihood.__call__(pinned host theta) -> host ll (vectorised getLikelihood)"},
                "gpu_launches": launches,
                "roofline": roof,
                "stage_ms_per_launch": stage_ms,
                "stage_rooflines": stage_roof,
                "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{sample_n} walkers of the same ensemble, literal per-walker oracle path, "
                                           f"Pool({cores}), {cpu_dt:.1f} s"},
                "parity_max_abs_dll_vs_cpu_sample": parity}
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
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
