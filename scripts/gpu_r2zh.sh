#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/yconv_loop.bin > gpurun_out/yconv_loop.log 2>&1; cat gpurun_out/yconv_loop.log
