#!/bin/bash
mkdir -p gpurun_out
for fw in 12 8; do
JX_K3W_FW=$fw timeout 200 python scripts/k3_phase_clocks.py > gpurun_out/k3w_clocks_fw$fw.log 2>&1; echo "== fw $fw"; tail -9 gpurun_out/k3w_clocks_fw$fw.log
done
