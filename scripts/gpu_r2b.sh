#!/bin/bash
# K3W (two walkers in flight) against K3: parity first (short timeout: a barrier bug would hang), then timing
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/smoke_ws.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_ws.log
tail -3 gpurun_out/smoke_ws.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_missing_data.py tests/test_calc_integ.py tests/test_sampler.py -m gpu -x -q > gpurun_out/pytest_ws.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_ws.log
tail -5 gpurun_out/pytest_ws.log
for ws in 1 0; do
  JX_K3_WS=$ws timeout 600 python bench.py --no-secondary --steps 5 > gpurun_out/bench_ws$ws.log 2> gpurun_out/bench_ws$ws.err; echo "bench ws=$ws rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ws$ws.log").read().strip().splitlines()[-1])
    print("ws=$ws", d["value"], d["ms_per_step"], d["stage_ms_per_launch"], d["roofline"]["frac"], d["parity_max_abs_dll_vs_cpu_sample"], d["state_checksum"])
except Exception as e:
    print("parse failed", e)
PY
done
timeout 300 python scripts/k3_phase_clocks.py > gpurun_out/k3w_clocks.log 2>&1; cat gpurun_out/k3w_clocks.log | tail -14
