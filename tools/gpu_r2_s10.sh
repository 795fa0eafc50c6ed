#!/bin/bash
mkdir -p gpurun_out
for ns in 2 3 4; do
python bench.py --cells 2000 --snvs 50000 --steps 5 --warmup 2 --no-cpu-baseline --no-gemm-leg --no-small-leg --rep-sub $ns > gpurun_out/r2_bench_rep$ns.json 2> gpurun_out/r2_bench_rep$ns.err; echo "rc=$?"
python - $ns <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_bench_rep%s.json" % sys.argv[1]))
r = d["replicate_sweep"]
print("n_sub", sys.argv[1], "device", r["ms_per_step"], "e2e", r["e2e"]["ms_per_step"], r["e2e"]["value"], "host", r["e2e"]["host_wall_ms_per_step"], r["e2e"]["rows_equal_device_leg"])
PY
done
