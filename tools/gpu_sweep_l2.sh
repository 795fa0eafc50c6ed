#!/bin/bash
# sweep the L2 blocking knobs of the placement GEMM at the full bench size
mkdir -p gpurun_out; : > gpurun_out/sweep_l2.log
for cfg in "20 1 32" "20 0 32" "12 1 32" "30 1 32" "40 1 32" "20 1 16" "20 1 64" "16 1 32"; do
  set -- $cfg
  echo "=== SLAB_MB=$1 HINTS=$2 RUN_LEN=$3" >> gpurun_out/sweep_l2.log
  SCS_SLAB_MB=$1 SCS_L2_HINTS=$2 SCS_RUN_LEN=$3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms_step %.1f pass1 %.1f pass2 %.1f frac %.3f sb_width %d clocks %s' % (d['ms_per_step'], r['pass1_ms'], r['pass2_ms'], r['frac'], d['config']['tiling']['superblock_width'], d['clocks']))" >> gpurun_out/sweep_l2.log 2>&1
done
cat gpurun_out/sweep_l2.log
