#!/bin/bash
# K3L2 phase B timing experiments: without its result stores / without its tile fetches / without both
mkdir -p gpurun_out
for tag in _b0 _b1 _b2 _b3; do for wl in synth255 synth511; do
  JX_CLK_TAG=$tag JX_CLK_WORKLOAD=$wl timeout 100 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_${wl}_zn$tag.log 2>&1
  echo "== $wl $tag"; tail -5 gpurun_out/k3l2_clocks_${wl}_zn$tag.log | tr '\n' ' '; echo
done; done
