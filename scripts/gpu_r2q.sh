#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_q.log
timeout 600 python bench.py --no-secondary --steps 10 > gpurun_out/bench_q.log 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_q.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "state_checksum", "parity_max_abs_dll_vs_cpu_sample")}, d["block_ms_per_step"])
print(d["stage_ms_per_launch"], d["roofline"]["frac"])
PY
timeout 300 python scripts/k3_phase_clocks.py > gpurun_out/k3w_clocks_q.log 2>&1; tail -10 gpurun_out/k3w_clocks_q.log
