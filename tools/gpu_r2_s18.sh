#!/bin/bash
# where does the interval path (variant 4 + fused pre-pass) overtake the single-chunk GEMM?  100k SNVs, one PT test
mkdir -p gpurun_out
for cells in 60 120 250 400; do
for algo in gemm interval; do
python bench.py --cells $cells --snvs 100000 --steps 20 --warmup 5 --algo $algo --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg --no-small-leg > gpurun_out/r2_cross_${cells}_$algo.json 2> gpurun_out/r2_cross.err || tail -3 gpurun_out/r2_cross.err
python - $cells $algo <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_cross_%s_%s.json" % (sys.argv[1], sys.argv[2])))
r = d["roofline"]
print("cells", sys.argv[1], sys.argv[2], "ms/step %.3f" % d["ms_per_step"], {k: round(r[k], 3) for k in ("ms_per_launch", "prepass_ms", "reduce_ms", "pass1_ms", "pass2_ms", "merge_ms") if r.get(k) is not None}, d["config"]["tiling"]["interval_threads"], d["config"]["tiling"]["interval_grid"])
PY
done
done
