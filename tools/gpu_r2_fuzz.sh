#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fuzz_parity.py 2 100 > gpurun_out/r2_fuzz.log 2>&1; echo rc=$?; tail -25 gpurun_out/r2_fuzz.log | cut -c1-500
