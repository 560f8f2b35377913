#!/bin/bash
mkdir -p gpurun_out
for tag in _ts1 _ts2; do
  JX_CLK_TAG=$tag JX_CLK_WORKLOAD=synth255 timeout 100 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_synth255_zo$tag.log 2>&1
  echo "== synth255 $tag"; tail -5 gpurun_out/k3l2_clocks_synth255_zo$tag.log | tr '\n' ' '; echo
done
