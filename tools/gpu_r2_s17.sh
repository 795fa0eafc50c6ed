#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "small_tree or sweep or batch or stream or two_ranks or streaming or chunk" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest17.log | cut -c1-300
python tools/sweep_stream_trace.py 4 > gpurun_out/r2_stream_trace6.log 2>&1; echo rc=$?; grep -v "ms  " gpurun_out/r2_stream_trace6.log | tail -12
for ns in 4 5 6; do
python bench.py --cells 2000 --snvs 50000 --steps 5 --warmup 2 --no-cpu-baseline --no-gemm-leg --no-small-leg --rep-sub $ns > gpurun_out/r2_bench_rep$ns.json 2> gpurun_out/r2_bench_rep$ns.err; echo "rc=$?"
python - $ns <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_bench_rep%s.json" % sys.argv[1]))
r = d["replicate_sweep"]
print("n_sub", sys.argv[1], "device", r["ms_per_step"], "e2e", r["e2e"]["ms_per_step"], r["e2e"]["value"], "host", r["e2e"]["host_wall_ms_per_step"], r["e2e"]["rows_equal_device_leg"], r["e2e"]["max_rel_dev_LR_vs_device_leg"])
PY
done
