#!/bin/bash
# pass 1 filling the count cache and the cached pass 2: one ncu --set full capture at 10k cells x
# 250k SNVs (same per-tile behaviour as 1M SNVs, a quarter of the replay time); SCS_PLACE_PASS2=recompute
# as $1 captures the plain pass 1 for comparison
set -u
MODE=${1:-cache}
mkdir -p gpurun_out
CMD="python bench.py --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0"
SCS_PLACE_PASS2=$MODE $CMD > gpurun_out/plain_${MODE}250.log 2>&1 &&
SCS_PLACE_PASS2=$MODE ncu --set full --clock-control none --import-source on -k "regex:scs_place_kernel|scs_colsum" -s 2 -c 2 -o gpurun_out/prof_${MODE}250 -f $CMD > gpurun_out/ncu_${MODE}250.log 2>&1
echo "$MODE capture rc=$?"
cut -c1-300 gpurun_out/plain_${MODE}250.log | tail -2
