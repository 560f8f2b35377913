#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3l2_szmap -s 2 -c 1 -f -o gpurun_out/r02f_k3l2_255_full python bench.py --workload synth255 --walkers 8192 --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_k3l2_f.log 2>&1; echo "ncu rc=$?"
