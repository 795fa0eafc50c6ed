#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "lrt or sweep or batch or tsv or stream or golden" > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest23.log | cut -c1-300
python bench.py --cells 2000 --snvs 50000 --steps 5 --warmup 2 --no-cpu-baseline --no-gemm-leg --no-small-leg --no-e2e > gpurun_out/r2_bench_rep.json 2> gpurun_out/r2_bench_rep.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_rep.json"))
r = d["replicate_sweep"]
print("device", r["ms_per_step"], r["pt_tests_per_s"], "placement", r["pass2_ms"], "lrt ok", r["lrt_converged"], "median it", r["median_newton_iterations"])
PY
