#!/bin/bash
# first GPU bring-up of the placement kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/bringup.log 2>&1
i=0
for args in "300 50" "1000 200" "3000 1000" "2000 300"; do
  i=$((i+1))
  echo "=== bringup $args" >> gpurun_out/bringup.log
  timeout 240 python tools/bringup_place.py $args >> gpurun_out/bringup.log 2>&1
  echo "exit $?" >> gpurun_out/bringup.log
done
echo "=== perf 100000 x 1000" >> gpurun_out/bringup.log
timeout 300 python tools/bringup_place.py 100000 1000 --perf >> gpurun_out/bringup.log 2>&1
echo "exit $?" >> gpurun_out/bringup.log
tail -60 gpurun_out/bringup.log
