#!/bin/bash
mkdir -p gpurun_out
run_n() {
  local n=$1; shift; local tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/scale2_n${n}${tag}.log 2> gpurun_out/scale2_n${n}${tag}.err
  echo "bench n=$n $tag rc=$?"
}
run_n 8 "" --no-secondary
run_n 8 _nccl --no-secondary --exchange nccl
run_n 4 "" --no-secondary
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale2_n*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["state_checksum"]["coords"], d["state_checksum"]["log_prob"], d["sampler"]["cuda_graph"], d["sampler"]["exchange"][:4], "e2e %.4g" % d["e2e"]["value"], {k: round(v, 4) for k, v in d["stage_ms_per_launch"].items()})
    except Exception as e:
        print(f, "parse failed", e)
PY
