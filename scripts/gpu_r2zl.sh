#!/bin/bash
# K3L2: memory-warp form of the y convolution (JX_K3L2_NT=257) against the shipped form, same box
mkdir -p gpurun_out
JX_K3L2_NT=257 timeout 400 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_zl.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_zl.log
for nt in 257 256; do for wl in synth255 synth511; do
  JX_K3L2_NT=$nt JX_CLK_WORKLOAD=$wl timeout 120 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_${wl}_zl$nt.log 2>&1
  echo "== $wl nt=$nt"; tail -5 gpurun_out/k3l2_clocks_${wl}_zl$nt.log | tr '\n' ' '; echo
done; done
