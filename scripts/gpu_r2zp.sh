#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_zp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_zp.log
timeout 200 python bench.py --workload synth255 --walkers 8192 --no-secondary --steps 3 > gpurun_out/bench_synth255_zp.log 2> gpurun_out/bench_synth255_zp.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_synth255_zp.log").read().strip().splitlines()[-1])
print("synth255 %.4g evals/s" % d["value"], "szmap %.3f ms" % d["stage_ms_per_launch"]["szmap"], "parity", d["parity_max_abs_dll_vs_cpu_sample"])
PY
