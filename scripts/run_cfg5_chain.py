#!/usr/bin/env python
"""BASELINE config 5 for real: the emcee chain schedule on the 511-pixel / 1024-point synthetic cluster with 65,536
walkers sharded over the GPUs of one box.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/run_cfg5_chain.py

Runs the reference's schedule (``joxsz_funcs.py:572-635``: preliminary rounds while the best log-probability improves,
burn-in, stored chain) with short counts, then `--steps` more timed iterations, and prints one JSON line (rank 0):
evals/s, ms per iteration, per-rank peak device memory, acceptance, parity of the final ensemble against the CPU
oracle on a sample, the state checksum (rank-count invariant), chain shapes.  The stored chain is sharded by rank.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402  (cluster builder, CPU pool, checksum helpers)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="synth511")
    ap.add_argument("--walkers", type=int, default=65536)
    ap.add_argument("--prefit", type=int, default=40)
    ap.add_argument("--nburn", type=int, default=40)
    ap.add_argument("--nsteps", type=int, default=120)
    ap.add_argument("--nthin", type=int, default=5)
    ap.add_argument("--steps", type=int, default=40, help="timed iterations after the schedule")
    ap.add_argument("--cpu-sample", type=int, default=32)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pool, cores = bench.make_pool() if rank == 0 else (None, 0)      # forked before CUDA is initialised
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.sampler import EnsembleSampler, mcmc_run
    from joxsz_b200.synthetic import FIDUCIAL

    bench.WORKLOAD = args.workload
    t0 = time.perf_counter()
    fit = bench.build_cluster(args.workload)
    W = args.walkers
    eng = BatchedLikelihood(fit, max_walkers=W // world + 64, device=local)
    setup_s = time.perf_counter() - t0
    mcmc = EnsembleSampler(W, eng.ndim, eng, seed=20260105, world_size=world, rank=rank,
                           group=(dist.group.WORLD if world > 1 else None), chain_shard=True)
    mcmc.initspread = 0.02
    fit.updateThawed([FIDUCIAL[n] for n in fit.thawed])
    np.random.seed(1234)                     # the initial ball is drawn on the host: same draws on every rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t1 = time.perf_counter()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()) as log:
        mcmc_run(mcmc, fit, args.nburn, args.nsteps, args.nthin, max_prefit=1, prefit_iterations=args.prefit)
    barrier()
    sched_s = time.perf_counter() - t1
    chain_shape = tuple(mcmc.get_chain().shape)
    # timed iterations on the device-resident ensemble
    blocks = bench.time_blocks(mcmc.step, args.steps, 3, barrier, reduce_max)
    ms = sorted(blocks)[1]
    chk = bench.state_checksum(mcmc)
    peak = torch.tensor([torch.cuda.max_memory_allocated() / 1e9], dtype=torch.float64, device="cuda")
    free, total = torch.cuda.mem_get_info()
    used = torch.tensor([(total - free) / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        peaks = [torch.zeros_like(peak) for _ in range(world)]
        useds = [torch.zeros_like(used) for _ in range(world)]
        dist.all_gather(peaks, peak)
        dist.all_gather(useds, used)
    else:
        peaks, useds = [peak], [used]
    if rank == 0:
        theta = mcmc.coords_host()[:args.cpu_sample]
        cpu_rate, cpu_dt, cpu_ll = bench.cpu_reference_rate(theta, pool, args.workload)
        pool.close()
        gpu_ll = eng(theta)
        lp_state = mcmc.log_prob_host()[:args.cpu_sample]
        fin = np.isfinite(cpu_ll)
        nsched = args.prefit + args.nburn + args.nsteps
        line = {"what": "BASELINE config 5: emcee chain schedule, synthetic cluster", "n_gpus": world,
                "config": bench.workload_config(W, world, args.workload),
                "schedule": {"prefit_iterations": args.prefit, "prefit_rounds": 1, "nburn": args.nburn,
                             "nsteps": args.nsteps, "nthin": args.nthin, "iterations": nsched,
                             "wall_s": sched_s, "evals_per_s_wall": W * nsched / sched_s,
                             "log": log.getvalue().strip().splitlines()},
                "timed": {"steps": args.steps, "blocks_ms_per_step": [b / args.steps for b in blocks],
                          "ms_per_step": ms / args.steps, "value": W * args.steps / (ms * 1e-3), "unit": "evals/s"},
                "cuda_graph": mcmc.graph_active, "graph_fallback_reason": mcmc._graph_failed,
                "acceptance_fraction": float(np.mean(mcmc.acceptance_fraction)),
                "chain_shape_per_rank": chain_shape, "chain_sharded": mcmc.chain_shard,
                "peak_torch_allocated_gb_per_rank": [float(p.item()) for p in peaks],
                "device_memory_in_use_gb_per_rank": [float(u.item()) for u in useds],
                "setup_s": setup_s, "state_checksum": chk,
                "parity": {"sample": int(len(theta)),
                           "max_abs_dll_gpu_vs_cpu": float(np.max(np.abs(gpu_ll[fin] - cpu_ll[fin]))) if fin.any() else None,
                           "max_abs_dll_chain_state_vs_cpu": float(np.max(np.abs(lp_state[fin] - cpu_ll[fin]))) if fin.any() else None,
                           "inf_mask_equal": bool(np.array_equal(fin, np.isfinite(gpu_ll)))},
                "cpu_baseline": {"value": cpu_rate, "unit": "evals/s", "cores": cores, "kind": "port",
                                 "sample": f"{len(theta)} walkers, literal per-walker oracle path, {cpu_dt:.1f} s"}}
        print(json.dumps(line), flush=True)
    mcmc.close()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
