mkdir -p gpurun_out
KRE="regex:scs_row|scs_missing|scs_expand|scs_place|scs_merge|scs_reduce|scs_lrt|scs_gt|scs_colsum|scs_permute"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/launches_interval.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list (interval) rc=$?"
ncu --set full --clock-control none --import-source on -k regex:scs_permute -s 2 -c 1 -o gpurun_out/prof_permute -f $CMD > gpurun_out/ncu_full_perm.log 2>&1
echo "full capture (permute) rc=$?"
