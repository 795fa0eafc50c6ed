#!/bin/bash
# one full ncu capture of the sweep leg's placement passes + LRT kernel (run under gpurun, 1 GPU)
set -u
mkdir -p gpurun_out
CMD="python tools/profile_sweep.py 2048"
timeout 600 $CMD > gpurun_out/sweep_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:scs_lrt_kernel|scs_place_kernel" -s 3 -c 3 -o gpurun_out/prof_sweep -f $CMD > gpurun_out/ncu_sweep.log 2>&1
echo "capture rc=$?"
tail -2 gpurun_out/sweep_plain.log | cut -c1-1500
ls -la gpurun_out/ | tail -5
