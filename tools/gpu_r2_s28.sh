#!/bin/bash
# lazy GEMM buffers + 2-CTA kernel removed: the whole GPU suite and the smoke run
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest28.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest28.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke28.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke28.log
