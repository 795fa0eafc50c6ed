#!/bin/bash
# round bench: reference arm, our arm, then ncu launch list + full capture of the GEMM
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:scs_ -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scs_place_kernel -s 2 -c 2 -o gpurun_out/prof_place_radix -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
CMD2="python bench.py --cells 1000 --snvs 100000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scs_place_kernel -s 2 -c 2 -o gpurun_out/prof_place_radix_1k -f $CMD2 > gpurun_out/ncu_full_1k.log 2>&1
echo "full capture 1k rc=$?"
cut -c1-400 gpurun_out/bench_default.json; cut -c1-300 gpurun_out/bench_reference.json
