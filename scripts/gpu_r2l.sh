#!/bin/bash
# K3L2 thread-count variants at both large-map sizes; then the whole GPU test suite
mkdir -p gpurun_out
for nt in 512 384 256; do for wl in synth255 synth511; do
  JX_K3L2=2 JX_K3L2_NT=$nt timeout 300 python bench.py --workload $wl --walkers 8192 --no-secondary --steps 3 > gpurun_out/bench_${wl}_nt$nt.log 2> gpurun_out/bench_${wl}_nt$nt.err; echo "bench $wl nt=$nt rc=$?"
done; done
python - <<'PY'
import json
for nt in (512, 384, 256):
    for wl in ("synth255", "synth511"):
        try:
            d = json.loads(open(f"gpurun_out/bench_{wl}_nt{nt}.log").read().strip().splitlines()[-1])
            print(wl, "nt", nt, "%.4g evals/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], "szmap %.3f ms" % d["stage_ms_per_launch"]["szmap"], "parity", d["parity_max_abs_dll_vs_cpu_sample"])
        except Exception as e:
            print(wl, nt, "parse failed", e)
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
