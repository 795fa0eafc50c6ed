#!/bin/bash
mkdir -p gpurun_out
python tools/sweep_stream_trace.py 4 > gpurun_out/r2_stream_trace6.log 2>&1; echo rc=$?; tail -32 gpurun_out/r2_stream_trace6.log
