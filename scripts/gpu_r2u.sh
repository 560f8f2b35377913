#!/bin/bash
# K3W warp split: 12 transform + 4 convolution warps against 8 + 8 (bit-identical results expected)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_u.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_u.log
for fw in 12 8; do
  JX_K3W_FW=$fw timeout 300 python bench.py --no-secondary --steps 10 > gpurun_out/bench_fw$fw.log 2> gpurun_out/bench_fw$fw.err; echo "bench fw=$fw rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_fw$fw.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "state_checksum", "parity_max_abs_dll_vs_cpu_sample")}, d["block_ms_per_step"])
print(d["stage_ms_per_launch"], d["roofline"]["frac"])
PY
done
