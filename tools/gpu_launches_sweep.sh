#!/bin/bash
# ncu launch list of the replicate-sweep leg (run under gpurun, 1 GPU)
set -u
mkdir -p gpurun_out
CMD="python tools/profile_sweep.py 2048"
timeout 600 $CMD > gpurun_out/sweep_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scs_row|scs_missing|scs_expand|scs_place|scs_merge|scs_reduce|scs_lrt|scs_gt" -c 200 --csv --log-file gpurun_out/sweep_launches.csv $CMD > gpurun_out/ncu_sweep_list.log 2>&1
echo "launch list rc=$?"
tail -1 gpurun_out/sweep_plain.log | cut -c1-600
