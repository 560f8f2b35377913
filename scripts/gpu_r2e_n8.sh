#!/bin/bash
# 8 GPUs of one box: rank-invariance tests (2 ranks, NCCL and peer-memory exchange, eager and graph), the scaling
# series N = 1, 2, 4, 8 of the default bench line, the NCCL-exchange variant at N = 8, and BASELINE config 5 as a chain
mkdir -p gpurun_out
nvidia-smi -L | head -8
run_n() {  # run_n N extra-args... -> gpurun_out/scale_nN<tag>.log
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scale_n${n}${tag}.log 2> gpurun_out/scale_n${n}${tag}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/scale_n${n}${tag}.log 2> gpurun_out/scale_n${n}${tag}.err
  fi
  echo "bench n=$n $tag rc=$?"
}
timeout 900 python -m pytest tests/test_sampler.py -m gpu -x -q -k "nccl or graph" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_nccl.log
tail -6 gpurun_out/pytest_nccl.log
run_n 8 "" 
run_n 8 _nccl --no-secondary --exchange nccl
run_n 8 _nograph --no-secondary --no-graph --exchange nccl
run_n 4 "" --no-secondary
run_n 2 "" --no-secondary
run_n 1 "" --no-secondary
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_n*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["state_checksum"]["coords"], d["sampler"]["cuda_graph"], d["sampler"]["exchange"][:4], "e2e %.4g" % d["e2e"]["value"])
        for k, v in (d.get("secondary") or {}).items():
            print("   ", k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "parity_max_abs_dll_vs_cpu_sample", "error", "peak_memory_gb")})
    except Exception as e:
        print(f, "parse failed", e)
PY
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 scripts/run_cfg5_chain.py > gpurun_out/cfg5_n8.log 2> gpurun_out/cfg5_n8.err; echo "cfg5 n8 rc=$?"
tail -c 1200 gpurun_out/cfg5_n8.err; tail -c 3000 gpurun_out/cfg5_n8.log
