#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "whole_path_on_random" -rf > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_pytest26.log | cut -c1-300
