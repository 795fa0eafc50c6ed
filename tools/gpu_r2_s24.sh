#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "lrt or golden or tsv or sweep or batch" > gpurun_out/r2_pytest24.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest24.log | cut -c1-400
