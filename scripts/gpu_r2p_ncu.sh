#!/bin/bash
# ncu --set full of the two DMMA GEMMs with their two-CTA tilings, and of K3L2 at 255 pixels (final code)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_dgemm|k7_filter" -s 4 -c 2 -f -o gpurun_out/r02d_k2_k7_full python bench.py --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_k27.log 2>&1; echo "ncu k2/k7 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3l2_szmap -s 2 -c 1 -f -o gpurun_out/r02d_k3l2_255_full python bench.py --workload synth255 --walkers 8192 --no-secondary --steps 2 --warmup 1 > gpurun_out/ncu_k3l2.log 2>&1; echo "ncu k3l2 rc=$?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
