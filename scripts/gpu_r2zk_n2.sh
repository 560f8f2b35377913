#!/bin/bash
# two GPUs of one box: the default bench line under torchrun (peer-memory exchange, graph, secondary block)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 > gpurun_out/bench_n2_final.log 2> gpurun_out/bench_n2_final.err; echo "bench n=2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2_final.log").read().strip().splitlines()[-1])
print(d["n_gpus"], "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["state_checksum"], d["sampler"]["cuda_graph"], d["sampler"]["exchange"], "e2e %.4g" % d["e2e"]["value"])
for k, v in (d.get("secondary") or {}).items():
    print("   ", k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "parity_max_abs_dll_vs_cpu_sample", "error")})
PY
tail -3 gpurun_out/bench_n2_final.err
