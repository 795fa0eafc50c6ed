#!/bin/bash
# round 2, session 4: tests (sweep engine), bench with the replicate e2e leg
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf -x > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest4.log | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-gemm-leg > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench4.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
r = d["replicate_sweep"]
print({k: r[k] for k in r if k not in ("tiling",)})
PY
