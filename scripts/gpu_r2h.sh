#!/bin/bash
mkdir -p gpurun_out
for wl in synth255 synth511; do
  JX_K3L2=1 timeout 300 python bench.py --workload $wl --walkers 8192 --no-secondary --steps 3 > gpurun_out/bench_${wl}_l21.log 2> gpurun_out/bench_${wl}_l21.err; echo "bench $wl rc=$?"
done
python - <<'PY'
import json
for wl in ("synth255", "synth511"):
    d = json.loads(open(f"gpurun_out/bench_{wl}_l21.log").read().strip().splitlines()[-1])
    print(wl, "%.4g evals/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], "szmap %.3f ms" % d["stage_ms_per_launch"]["szmap"], "parity", d["parity_max_abs_dll_vs_cpu_sample"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3l2_szmap -s 2 -c 1 -f -o gpurun_out/r02b_k3l2_255_full python bench.py --workload synth255 --walkers 8192 --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_k3l2.log 2>&1; echo "ncu rc=$?"
