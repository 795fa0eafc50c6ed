#!/bin/bash
# ncu launch list of the 1,000 cells x 100k SNVs configuration (run under gpurun, 1 GPU)
set -u
mkdir -p gpurun_out
CMD="python bench.py --cells 1000 --snvs 100000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --replicates 0"
timeout 600 $CMD > gpurun_out/plain_1k.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scs_row|scs_missing|scs_expand|scs_place|scs_merge|scs_reduce|scs_lrt|scs_gt" -c 200 --csv --log-file gpurun_out/launches_1k.csv $CMD > gpurun_out/ncu_list_1k.log 2>&1
echo "launch list rc=$?"
tail -1 gpurun_out/plain_1k.log | cut -c1-300
