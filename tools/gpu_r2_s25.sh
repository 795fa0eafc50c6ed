#!/bin/bash
mkdir -p gpurun_out
CMD2="python bench.py --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg --no-small-leg"
$CMD2 > gpurun_out/r2e_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_place_iv4|scs_iv4_leaf|scs_permute_reduce|scs_lrt_tree" -s 4 -c 4 -o gpurun_out/r2e_prof_10k_250k -f $CMD2 > gpurun_out/r2e_ncu_full2.log 2>&1
echo "full capture rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg --no-small-leg"
$CMD > gpurun_out/r2e_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scs_" -c 200 --csv --log-file gpurun_out/r2e_launches_10k.csv $CMD > gpurun_out/r2e_ncu_list1.log 2>&1
echo "launch list rc=$?"
