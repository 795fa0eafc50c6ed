#!/bin/bash
# ncu launch list + one full capture of the placement GEMM (run under gpurun, 1 GPU)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:scs_ -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scs_place_kernel -s 2 -c 2 -o gpurun_out/prof_place -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/plain.log | cut -c1-300
ls -la gpurun_out/
