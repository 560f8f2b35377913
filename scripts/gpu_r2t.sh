#!/bin/bash
# phase clocks of the large-map kernel K3L2 at both sizes
mkdir -p gpurun_out
for wl in synth255 synth511; do
  JX_CLK_WORKLOAD=$wl timeout 300 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_$wl.log 2>&1; echo "== $wl"; tail -7 gpurun_out/k3l2_clocks_$wl.log
done
