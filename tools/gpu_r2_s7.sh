#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "lrt" > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest7.log | cut -c1-400
timeout 600 python tools/lrt_timing.py 1000 10000 > gpurun_out/r2_lrt_timing.log 2>&1; echo "timing rc=$?"; cat gpurun_out/r2_lrt_timing.log | cut -c1-600 | tail -20
