#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/batch_invariance_check.py 2>&1 | tee gpurun_out/batch_invariance.log | tail -40
for ws in 0; do JX_K3_WS=$ws timeout 600 python scripts/batch_invariance_check.py 2>&1 | tail -12 > gpurun_out/batch_invariance_ws$ws.log; echo "--- JX_K3_WS=$ws"; cat gpurun_out/batch_invariance_ws$ws.log; done
