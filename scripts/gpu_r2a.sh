#!/bin/bash
# round 2, first GPU pass: GPU tests, then the default bench line (headline + secondary block)
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r2a.log 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r2a.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2a.log").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "block_ms_per_step", "state_checksum", "sampler", "parity_max_abs_dll_vs_cpu_sample")})
    print("e2e", d["e2e"]["value"], "numpy", d["e2e_numpy"] and d["e2e_numpy"]["value"])
    print("roof", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["stage_ms_per_launch"])
    for k, v in d.get("secondary", {}).items():
        print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "stage_ms_per_launch", "parity_max_abs_dll_vs_cpu_sample", "error", "setup_s", "peak_memory_gb")})
except Exception as e:
    print("parse failed", e)
PY
