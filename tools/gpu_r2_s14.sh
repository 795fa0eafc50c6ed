#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --cells 2000 --snvs 20000 --steps 2 --warmup 1 --no-cpu-baseline --no-gemm-leg --no-small-leg --no-e2e --replicates 256"
$CMD > gpurun_out/r2_plain_iv5.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:scs_place_iv5" -c 1 -o gpurun_out/r2_prof_iv5 -f $CMD > gpurun_out/r2_ncu_iv5.log 2>&1
echo "ncu rc=$?"
