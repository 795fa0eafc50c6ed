#!/bin/bash
# round evidence (run under gpurun, 1 GPU): GPU tests, reference arm, our arm, ncu launch lists and
# full captures of the dominant kernels of both placement algorithms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
KRE="regex:scs_row|scs_missing|scs_expand|scs_place|scs_merge|scs_reduce|scs_lrt|scs_gt|scs_colsum|scs_permute"
# --- 10k cells x 1M SNVs, interval path (default)
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/launches_interval.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list (interval) rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scs_place_interval -s 2 -c 1 -o gpurun_out/prof_interval -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture (interval) rc=$?"
# --- the same workload through the GEMM (count cache): launch list; full capture at 250k SNVs (ncu would have to save and
# restore the 80 GB count cache around every replay pass at 1M)
CMDG="python bench.py --algo gemm --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0"
$CMDG > gpurun_out/plain_gemm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 200 --csv --log-file gpurun_out/launches_gemm.csv $CMDG > gpurun_out/ncu_list_gemm.log 2>&1
echo "launch list (gemm) rc=$?"
CMDG2="python bench.py --algo gemm --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0"
$CMDG2 > gpurun_out/plain_gemm250.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_place_kernel|scs_colsum" -s 2 -c 2 -o gpurun_out/prof_gemm_cache250 -f $CMDG2 > gpurun_out/ncu_full_gemm.log 2>&1
echo "full capture (gemm) rc=$?"
cut -c1-600 gpurun_out/bench_default.json; echo; cut -c1-300 gpurun_out/bench_reference.json
