#!/bin/bash
# round 2 evidence (run under gpurun, 1 GPU): launch lists of the bench step and of the replicate leg, full captures of
# the dominant kernels (placement iv4 + pre-pass at 250k SNVs, the cluster LRT kernel, the small-tree placement iv5)
mkdir -p gpurun_out
KRE="regex:scs_"
# --- 10k cells x 1M SNVs (default = interval path): launch list of two steps
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg --no-small-leg"
$CMD > gpurun_out/r2e_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/r2e_launches_10k.csv $CMD > gpurun_out/r2e_ncu_list1.log 2>&1
echo "launch list (10k x 1M) rc=$?"
# --- full capture at 250k SNVs: pre-pass, placement, LRT
CMD2="python bench.py --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg --no-small-leg"
$CMD2 > gpurun_out/r2e_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_place_iv4|scs_iv4_leaf|scs_permute_reduce|scs_lrt_tree" -s 4 -c 4 -o gpurun_out/r2e_prof_10k_250k -f $CMD2 > gpurun_out/r2e_ncu_full2.log 2>&1
echo "full capture (iv4 + prepass + lrt) rc=$?"
# --- replicate sweep leg (1250 x 100 cells x 10k SNVs): launch list, full capture of the placement and the LRT kernels
CMD3="python bench.py --cells 2000 --snvs 20000 --steps 2 --warmup 1 --no-cpu-baseline --no-gemm-leg --no-small-leg --no-e2e"
$CMD3 > gpurun_out/r2e_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/r2e_launches_sweep.csv $CMD3 > gpurun_out/r2e_ncu_list3.log 2>&1
echo "launch list (sweep) rc=$?"
CMD4="python bench.py --cells 2000 --snvs 20000 --steps 2 --warmup 1 --no-cpu-baseline --no-gemm-leg --no-small-leg --no-e2e --replicates 256"
$CMD4 > gpurun_out/r2e_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_place_iv5|scs_lrt_kernel|scs_gt_reduce" -c 3 -o gpurun_out/r2e_prof_sweep256 -f $CMD4 > gpurun_out/r2e_ncu_full4.log 2>&1
echo "full capture (sweep) rc=$?"
# --- 1k cells x 100k SNVs (configs[2]): launch list
CMD5="python bench.py --cells 1000 --snvs 100000 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg --no-small-leg"
$CMD5 > gpurun_out/r2e_plain5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/r2e_launches_1k.csv $CMD5 > gpurun_out/r2e_ncu_list5.log 2>&1
echo "launch list (1k x 100k) rc=$?"
