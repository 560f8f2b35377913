#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r2m.log 2> gpurun_out/bench_r2m.err; echo "bench rc=$?"
for w in 8192 16384; do
  timeout 300 python bench.py --walkers $w --no-secondary --steps 50 > gpurun_out/bench_w$w.log 2> gpurun_out/bench_w$w.err; echo "bench W=$w rc=$?"
  timeout 300 python bench.py --walkers $w --no-secondary --steps 50 --no-graph > gpurun_out/bench_w${w}_nograph.log 2> gpurun_out/bench_w${w}_nograph.err; echo "bench W=$w nograph rc=$?"
done
python - <<'PY'
import json
for f in ("bench_r2m", "bench_w8192", "bench_w8192_nograph", "bench_w16384", "bench_w16384_nograph"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        st = d["stage_ms_per_launch"]
        lk = sum(v for k, v in st.items() if k != "xray")
        print(f, "W", d["config"]["walkers"], "%.4g evals/s" % d["value"], "step %.4f ms" % d["ms_per_step"], "2 x likelihood kernels %.4f" % (2 * lk), "rest %.4f ms" % (d["ms_per_step"] - 2 * lk), {k: round(v, 4) for k, v in st.items()}, "roof", round(d["roofline"]["frac"], 3), d["roofline"]["peak"])
    except Exception as e:
        print(f, "parse failed", e)
PY
