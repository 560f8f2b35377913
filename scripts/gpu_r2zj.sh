#!/bin/bash
mkdir -p gpurun_out
for tag in "" _nostg; do
  JX_CLK_TAG=$tag timeout 200 python scripts/k3_phase_clocks.py > gpurun_out/k3w_clocks_zj$tag.log 2>&1; echo "== k3w $tag"; tail -9 gpurun_out/k3w_clocks_zj$tag.log
done
