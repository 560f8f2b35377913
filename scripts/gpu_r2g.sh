#!/bin/bash
# K3L2 (two CTAs per SM) against K3L: parity tests of the large maps, then the two synthetic workloads both ways
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_large.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_large.log
for v in 1 0; do for wl in synth255 synth511; do
  JX_K3L2=$v timeout 300 python bench.py --workload $wl --walkers 8192 --no-secondary --steps 3 > gpurun_out/bench_${wl}_l2$v.log 2> gpurun_out/bench_${wl}_l2$v.err; echo "bench $wl k3l2=$v rc=$?"
done; done
python - <<'PY'
import json
for v in (1, 0):
    for wl in ("synth255", "synth511"):
        try:
            d = json.loads(open(f"gpurun_out/bench_{wl}_l2{v}.log").read().strip().splitlines()[-1])
            print(wl, "k3l2", v, "%.4g evals/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], "szmap %.3f ms" % d["stage_ms_per_launch"]["szmap"], "parity", d["parity_max_abs_dll_vs_cpu_sample"])
        except Exception as e:
            print(wl, v, "parse failed", e)
PY
