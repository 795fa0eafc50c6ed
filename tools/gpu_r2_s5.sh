#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "sweep or batch or two_ranks" > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest5.log | cut -c1-300
python tools/sweep_timeline.py 1250 4 > gpurun_out/r2_sweep_timeline4.log 2>&1; tail -8 gpurun_out/r2_sweep_timeline4.log
python tools/sweep_timeline.py 1250 8 > gpurun_out/r2_sweep_timeline8.log 2>&1; tail -4 gpurun_out/r2_sweep_timeline8.log
