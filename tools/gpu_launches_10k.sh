#!/bin/bash
# ncu launch list of the 10k cells x 1M SNVs configuration (run under gpurun, 1 GPU)
set -u
mkdir -p gpurun_out
KRE="regex:scs_row|scs_missing|scs_expand|scs_place|scs_merge|scs_reduce|scs_lrt|scs_gt"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
tail -1 gpurun_out/plain.log | cut -c1-300
