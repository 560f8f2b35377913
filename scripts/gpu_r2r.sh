#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3w_szmap -s 2 -c 1 -f -o gpurun_out/r02e_k3w_full python bench.py --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_k3w.log 2>&1; echo "ncu k3w rc=$?"
timeout 300 python scripts/k3_phase_clocks.py > gpurun_out/k3w_clocks.log 2>&1; tail -10 gpurun_out/k3w_clocks.log
