#!/bin/bash
# the driver's round-end sequence on one GPU at the final code, then ncu of the large-map kernel at both sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
r = json.loads(open("gpurun_out/bench_ref.log").read().strip().splitlines()[-1])
d = json.loads(open("gpurun_out/bench_final.log").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"])
print({k: d[k] for k in ("value", "ms_per_step", "state_checksum", "parity_max_abs_dll_vs_cpu_sample")})
print("e2e", d["e2e"]["value"], "numpy", d["e2e_numpy"]["value"], "collapsed", d["collapsed_mode"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["peak"], d["stage_ms_per_launch"])
for k, v in d["secondary"].items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "parity_max_abs_dll_vs_cpu_sample", "error", "map_kernel")}, v.get("stage_ms_per_launch", {}).get("szmap"))
PY
