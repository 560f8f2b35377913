#!/bin/bash
# K3L2 final of this pass: tests, phase clocks, bench lines of the two large-map workloads
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_zc.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_zc.log
for wl in synth255 synth511; do
  JX_CLK_WORKLOAD=$wl timeout 120 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_${wl}_zc.log 2>&1
  echo "== $wl"; tail -5 gpurun_out/k3l2_clocks_${wl}_zc.log | tr '\n' ' '; echo
done
for wl in synth255 synth511; do
  timeout 300 python bench.py --workload $wl --walkers 8192 --no-secondary --steps 3 > gpurun_out/bench_${wl}_zc.log 2> gpurun_out/bench_${wl}_zc.err; echo "bench $wl rc=$?"
done
python - <<'PY'
import json
for wl in ("synth255", "synth511"):
    d = json.loads(open(f"gpurun_out/bench_{wl}_zc.log").read().strip().splitlines()[-1])
    print(wl, "%.4g evals/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], "szmap %.3f ms" % d["stage_ms_per_launch"]["szmap"], "parity", d["parity_max_abs_dll_vs_cpu_sample"], d["stage_ms_per_launch"])
PY
