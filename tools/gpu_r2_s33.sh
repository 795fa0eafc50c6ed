#!/bin/bash
# per-cell kernel v2 (512 threads, clade rows in registers, interval table in shared memory): tests, fuzz, bench leg, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "per_cell" > gpurun_out/r2_pytest33a.log 2>&1; echo "percell pytest rc=$?"; tail -5 gpurun_out/r2_pytest31a.log
timeout 200 python tools/fuzz_parity.py 9 30 percell > gpurun_out/r2_fuzz_percell.log 2>&1; echo "fuzz rc=$?"; tail -5 gpurun_out/r2_fuzz_percell.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-gemm-leg --no-small-leg --replicates 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench33.json 2> gpurun_out/r2_bench33.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench33.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "launches", d["gpu_launches"])
    print("per_cell_fn", d.get("per_cell_fn"))
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r2_bench33.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scs_place_percell -c 1 -o gpurun_out/r2f_prof_percell -f python bench.py --steps 1 --warmup 1 --snvs 100000 --no-gemm-leg --no-small-leg --replicates 0 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_percell.log 2>&1; echo "ncu rc=$?"
