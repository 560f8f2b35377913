#!/bin/bash
# default bench line (with the secondary block), then ncu: launch list, K3W full capture, K3L full capture at 255 px
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2d.log 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"
grep "^\[bench" gpurun_out/bench_r2d.err | tail -30
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2d.log").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "state_checksum", "parity_max_abs_dll_vs_cpu_sample")})
    print("e2e", d["e2e"]["value"], "numpy", d["e2e_numpy"] and d["e2e_numpy"]["value"], "roof", d["roofline"]["frac"], d["stage_ms_per_launch"])
    for k, v in d.get("secondary", {}).items():
        print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "stage_ms_per_launch", "parity_max_abs_dll_vs_cpu_sample", "error", "setup_s", "peak_memory_gb", "graph")})
except Exception as e:
    print("parse failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3w_szmap -s 2 -c 1 -f -o gpurun_out/r02a_k3w_full python bench.py --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_k3w.log 2>&1; echo "ncu k3w rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3l_szmap -s 2 -c 1 -f -o gpurun_out/r02a_k3l255_full python bench.py --workload synth255 --walkers 8192 --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_k3l.log 2>&1; echo "ncu k3l rc=$?"
timeout 300 compute-sanitizer --tool racecheck --kernel-regex kns=k3w python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck_k3w.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck_k3w.log
timeout 300 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck_smoke.log
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
