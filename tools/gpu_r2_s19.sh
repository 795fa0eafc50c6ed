#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "interval or prepare or streaming or placement or ten_thousand or full_baseline or tsv" > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest19.log | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gemm-leg --replicates 0 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench7.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench7.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print({k: d["roofline"].get(k) for k in ("ms_per_launch", "prepass_ms", "prepass_GBps", "reduce_ms")})
s = d["single_test_1k_cells"]; print("1k:", s["ms_per_step"], s["kernel_ms"], s["e2e"]["ms_per_step"])
PY
