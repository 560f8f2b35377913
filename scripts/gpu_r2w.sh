#!/bin/bash
# K3L2 with the line gathered once for all branches; prefetch depth and thread-count variants
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_w.log
for tag in _pf8 _pf16; do for nt in 384 256; do for wl in synth255 synth511; do
  JX_K3L2_NT=$nt JX_CLK_TAG=$tag JX_CLK_WORKLOAD=$wl timeout 120 python scripts/k3_phase_clocks.py 4096 > gpurun_out/k3l2_clocks_${wl}${tag}_nt$nt.log 2>&1
  echo "== $wl $tag nt=$nt"; tail -5 gpurun_out/k3l2_clocks_${wl}${tag}_nt$nt.log | tr '\n' ' '; echo
done; done; done
