#!/bin/bash
# experiment: the two warp groups of K3W alone (the other group only keeps the barrier protocol)
mkdir -p gpurun_out
for tag in _nom _nof; do
  JX_CLK_TAG=$tag timeout 200 python scripts/k3_phase_clocks.py > gpurun_out/k3w_clocks$tag.log 2>&1; echo "== $tag"; tail -9 gpurun_out/k3w_clocks$tag.log
done
