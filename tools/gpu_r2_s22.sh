#!/bin/bash
mkdir -p gpurun_out
python tools/sweep_stream_trace.py 4 2>&1 | grep "n_sub\|device stages\|per-sub" | cut -c1-300
