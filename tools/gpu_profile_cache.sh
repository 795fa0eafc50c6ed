#!/bin/bash
# pass 1 with and without the count-cache stores, and the cached pass 2: one ncu --set full capture
# each at 10k cells x 250k SNVs (same per-tile behaviour as 1M SNVs, a quarter of the replay time)
set -u
mkdir -p gpurun_out
CMD="python bench.py --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0"
SCS_PLACE_PASS2=cache $CMD > gpurun_out/plain_c250.log 2>&1 &&
SCS_PLACE_PASS2=cache ncu --set full --clock-control none --import-source on -k "regex:scs_place_kernel|scs_colsum" -s 2 -c 2 -o gpurun_out/prof_cache250 -f $CMD > gpurun_out/ncu_c250.log 2>&1
echo "cache capture rc=$?"
SCS_PLACE_PASS2=recompute ncu --set full --clock-control none --import-source on -k "regex:scs_place_kernel" -s 2 -c 1 -o gpurun_out/prof_recomp250 -f $CMD > gpurun_out/ncu_r250.log 2>&1
echo "recompute capture rc=$?"
cut -c1-300 gpurun_out/plain_c250.log | tail -2
