#!/bin/bash
# per-cell FN kernel: its tests, then the whole suite, then the bench line with the new leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "per_cell" > gpurun_out/r2_pytest29a.log 2>&1; echo "percell pytest rc=$?"; tail -15 gpurun_out/r2_pytest29a.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-gemm-leg --no-small-leg --replicates 0 --no-cpu-baseline > gpurun_out/r2_bench29.json 2> gpurun_out/r2_bench29.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench29.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    print("per_cell_fn", d.get("per_cell_fn"))
except Exception as e:
    print("no bench line", e)
PY
tail -5 gpurun_out/r2_bench29.err
