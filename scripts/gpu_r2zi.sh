#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_zi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_zi.log
for wl in synth255 synth511; do
  JX_CLK_WORKLOAD=$wl timeout 120 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_${wl}_zi.log 2>&1
  echo "== $wl"; tail -5 gpurun_out/k3l2_clocks_${wl}_zi.log | tr '\n' ' '; echo
done
