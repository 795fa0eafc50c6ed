#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/radix.log
for mode in radix planes; do
for args in "300 50" "1000 200" "600 3968 --stress" "400 4100 --stress" "300 9000 --stress" "300 10000" "2000 1000"; do
  echo "=== $mode $args" >> gpurun_out/radix.log
  SCS_PLACE_MODE=$mode timeout 300 python tools/bringup_place.py $args 2>&1 | grep -v "^ctx" >> gpurun_out/radix.log
done; done
for mode in radix planes; do
  echo "=== perf $mode 100000 x 1000" >> gpurun_out/radix.log
  SCS_PLACE_MODE=$mode timeout 300 python tools/bringup_place.py 100000 1000 --perf 2>&1 | tail -4 >> gpurun_out/radix.log
  echo "=== bench $mode" >> gpurun_out/radix.log
  SCS_PLACE_MODE=$mode python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms_step %.1f pass1 %.1f pass2 %.1f frac %.3f tiling %s LR %.4f clocks %s' % (d['ms_per_step'], r['pass1_ms'], r['pass2_ms'], r['frac'], d['config']['tiling'], d['pt_tests']['LR'][0], d['clocks']))" >> gpurun_out/radix.log 2>&1
done
cat gpurun_out/radix.log
