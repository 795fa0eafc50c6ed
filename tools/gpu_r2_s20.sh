#!/bin/bash
mkdir -p gpurun_out
CMD2="python bench.py --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg --no-small-leg"
$CMD2 > gpurun_out/r2e_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_permute_reduce" -s 1 -c 1 -o gpurun_out/r2_prof_prepass -f $CMD2 > gpurun_out/r2_ncu_prepass.log 2>&1
echo "rc=$?"
