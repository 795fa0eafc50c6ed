#!/bin/bash
# round evidence (run under gpurun, 1 GPU): reference arm, our arm, then per configuration a plain
# run followed by the ncu launch list and one full capture of the dominant kernels
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
KRE="regex:scs_row|scs_missing|scs_expand|scs_place|scs_merge|scs_reduce|scs_lrt|scs_gt"
# --- 10k cells x 1M SNVs
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scs_place_kernel -s 2 -c 2 -o gpurun_out/prof_place_radix -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
# --- 1k cells x 100k SNVs
CMD2="python bench.py --cells 1000 --snvs 100000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --replicates 0"
$CMD2 > gpurun_out/plain_1k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/launches_1k.csv $CMD2 > gpurun_out/ncu_list_1k.log 2>&1
echo "launch list 1k rc=$?"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_place_kernel|scs_lrt_kernel" -s 9 -c 3 -o gpurun_out/prof_place_radix_1k -f $CMD2 > gpurun_out/ncu_full_1k.log 2>&1
echo "full capture 1k rc=$?"
# --- replicate sweep, 2048 x (100 cells x 10k SNVs)
CMD3="python tools/profile_sweep.py 2048"
$CMD3 > gpurun_out/sweep_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/sweep_launches.csv $CMD3 > gpurun_out/ncu_sweep_list.log 2>&1
echo "launch list sweep rc=$?"
$CMD3 > gpurun_out/sweep_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_lrt_kernel|scs_place_kernel" -s 3 -c 3 -o gpurun_out/prof_sweep -f $CMD3 > gpurun_out/ncu_sweep.log 2>&1
echo "full capture sweep rc=$?"
cut -c1-400 gpurun_out/bench_default.json; cut -c1-300 gpurun_out/bench_reference.json; tail -1 gpurun_out/sweep_plain.log | cut -c1-300
