#!/bin/bash
# round 2, session 1 (gpurun, 1 GPU): GPU tests, default bench, launch list and a full capture of the new interval kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2_pytest1.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
KRE="regex:scs_"
CMD="python bench.py --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg"
$CMD > gpurun_out/r2_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/r2_launches_iv4_250k.csv $CMD > gpurun_out/r2_ncu_list1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_place_iv4|scs_permute_reduce" -s 2 -c 2 -o gpurun_out/r2_prof_iv4_250k -f $CMD > gpurun_out/r2_ncu_full1.log 2>&1
echo "full capture rc=$?"
cut -c1-1500 gpurun_out/r2_bench1.json; echo
tail -5 gpurun_out/r2_bench1.err
