#!/bin/bash
# 8 GPUs of one box: the default bench line at N = 8 (peer-memory exchange, graph, secondary block), the round-1
# configuration for comparison (NCCL all-gather, kernel-by-kernel launches), N = 1 on the same box, and BASELINE
# config 5 as a chain (65,536 walkers, 1024-point grid, 511-pixel map, the mcmc_run schedule)
mkdir -p gpurun_out
nvidia-smi -L | head -8
run_n() {
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/scale_n${n}${tag}.log 2> gpurun_out/scale_n${n}${tag}.err
  else
    timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/scale_n${n}${tag}.log 2> gpurun_out/scale_n${n}${tag}.err
  fi
  echo "bench n=$n $tag rc=$?"
}
run_n 8 ""
grep "^\[bench" gpurun_out/scale_n8.err | tail -6
run_n 8 _nccl_nograph --no-secondary --no-graph --exchange nccl
run_n 8 _nccl --no-secondary --exchange nccl
run_n 1 "" --no-secondary
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_n*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["state_checksum"]["coords"], d["state_checksum"]["log_prob"], d["sampler"]["cuda_graph"], d["sampler"]["exchange"][:4], d["sampler"].get("p2p_fallback_reason"), "e2e %.4g" % d["e2e"]["value"])
        for k, v in (d.get("secondary") or {}).items():
            print("   ", k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "parity_max_abs_dll_vs_cpu_sample", "error", "peak_memory_gb")})
    except Exception as e:
        print(f, "parse failed", e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 scripts/run_cfg5_chain.py > gpurun_out/cfg5_n8.log 2> gpurun_out/cfg5_n8.err; echo "cfg5 n8 rc=$?"
tail -c 600 gpurun_out/cfg5_n8.err; tail -c 3500 gpurun_out/cfg5_n8.log
