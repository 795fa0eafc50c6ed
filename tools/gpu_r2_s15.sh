#!/bin/bash
mkdir -p gpurun_out
python bench.py --cells 2000 --snvs 50000 --steps 5 --warmup 2 --no-cpu-baseline --no-gemm-leg --no-small-leg --no-e2e > gpurun_out/r2_bench_rep_iv5.json 2> gpurun_out/r2_bench_rep_iv5.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench_rep_iv5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_rep_iv5.json"))
r = d["replicate_sweep"]
print({k: r[k] for k in ("ms_per_step", "pt_tests_per_s", "pass2_ms")})
PY
