#!/bin/bash
mkdir -p gpurun_out
for ns in 12 8; do python tools/sweep_diff.py $ns > gpurun_out/r2_sweep_diff.log 2>&1; echo rc=$?; grep "one :\|sub0:\|differing" gpurun_out/r2_sweep_diff.log | cut -c1-600; done
