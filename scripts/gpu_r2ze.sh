#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_large_maps.py -m gpu -x -q > gpurun_out/pytest_ze.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_ze.log
