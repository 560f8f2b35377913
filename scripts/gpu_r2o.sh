#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_large_maps.py tests/test_calc_integ.py -m gpu -x -q > gpurun_out/pytest_k27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_k27.log
timeout 600 python bench.py --steps 10 > gpurun_out/bench_k27.log 2> gpurun_out/bench_k27.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_k27.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "state_checksum", "parity_max_abs_dll_vs_cpu_sample")})
print(d["stage_ms_per_launch"], {k: (v.get("frac")) for k, v in d["stage_rooflines"].items() if "frac" in v})
for k, v in d["secondary"].items():
    print(k, v.get("value"), v.get("stage_ms_per_launch"))
PY
