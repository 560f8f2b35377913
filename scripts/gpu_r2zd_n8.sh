#!/bin/bash
# BASELINE config 5 as a chain on 8 GPUs with the round's final large-map kernel
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 scripts/run_cfg5_chain.py > gpurun_out/cfg5_n8.log 2> gpurun_out/cfg5_n8.err; echo "cfg5 n8 rc=$?"
tail -c 600 gpurun_out/cfg5_n8.err; tail -c 3500 gpurun_out/cfg5_n8.log
