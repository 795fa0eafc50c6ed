#!/bin/bash
mkdir -p gpurun_out
python tools/sweep_stages.py 4 > gpurun_out/r2_sweep_stages.log 2>&1; echo rc=$?; tail -6 gpurun_out/r2_sweep_stages.log
