#!/bin/bash
# K3L2: loads in flight per gather chunk of a row transform
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_zf.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_zf.log
for tag in _g8 _g16 _g32; do for wl in synth255 synth511; do
  JX_CLK_TAG=$tag JX_CLK_WORKLOAD=$wl timeout 120 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_${wl}${tag}.log 2>&1
  echo "== $wl $tag"; tail -5 gpurun_out/k3l2_clocks_${wl}${tag}.log | tr '\n' ' '; echo
done; done
