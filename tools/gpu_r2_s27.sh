#!/bin/bash
mkdir -p gpurun_out
export SCS_IV4_LEAF_SPLIT=1
timeout 1500 python -m pytest tests -m gpu -q -x -k "interval or ten_thousand or full_baseline or placement" > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest27.log | cut -c1-300
for split in 0 1; do
SCS_IV4_LEAF_SPLIT=$split python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gemm-leg --replicates 0 --no-e2e --no-small-leg > gpurun_out/r2_bench_split$split.json 2> gpurun_out/r2_bench_split.err; echo "bench rc=$?"
python - $split <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_bench_split%s.json" % sys.argv[1]))
print("split", sys.argv[1], "ms/step", d["ms_per_step"], {k: d["roofline"].get(k) for k in ("ms_per_launch", "prepass_ms", "reduce_ms")}, "LR", d["pt_tests"]["LR"][0])
PY
done
