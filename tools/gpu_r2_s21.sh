#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "small_tree or sweep or batch or stream or two_ranks" > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest21.log | cut -c1-300
python bench.py --cells 2000 --snvs 50000 --steps 5 --warmup 2 --no-cpu-baseline --no-gemm-leg --no-small-leg > gpurun_out/r2_bench_rep.json 2> gpurun_out/r2_bench_rep.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_rep.json"))
r = d["replicate_sweep"]
print("device", r["ms_per_step"], r["pt_tests_per_s"], "placement", r["pass2_ms"], "e2e", r["e2e"]["ms_per_step"], r["e2e"]["value"], r["e2e"]["rows_equal_device_leg"])
PY
python tools/sweep_stream_trace.py 4 2>&1 | grep "device stages" | cut -c1-300
