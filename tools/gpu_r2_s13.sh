#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "small_tree or sweep or batch or stream or two_ranks" > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_pytest13.log | cut -c1-400
python bench.py --cells 2000 --snvs 50000 --steps 5 --warmup 2 --no-cpu-baseline --no-gemm-leg --no-small-leg > gpurun_out/r2_bench_rep_iv5.json 2> gpurun_out/r2_bench_rep_iv5.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench_rep_iv5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_rep_iv5.json"))
r = d["replicate_sweep"]
print({k: r[k] for k in r if k not in ("tiling", "e2e")}); print(r["e2e"]); print(r["tiling"])
PY
