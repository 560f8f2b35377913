#!/bin/bash
# K3L2: y convolution on shared-memory tiles streamed with cp.async
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_za.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_za.log
for wl in synth255 synth511; do
  JX_CLK_WORKLOAD=$wl timeout 120 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_${wl}_cpasync.log 2>&1
  echo "== $wl"; tail -5 gpurun_out/k3l2_clocks_${wl}_cpasync.log | tr '\n' ' '; echo
done
