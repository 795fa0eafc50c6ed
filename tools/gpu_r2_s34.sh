#!/bin/bash
# final code of the round (per-cell kernel v2 with odd position blocks): whole GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest34.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest34.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke34.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke34.log
timeout 600 python bench.py > gpurun_out/r2_bench34.json 2> gpurun_out/r2_bench34.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench34.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch", "prepass_ms")})
    print("per_cell_fn", d.get("per_cell_fn"))
    print("small", d["single_test_1k_cells"]["ms_per_step"], "rep", d["replicate_sweep"]["ms_per_step"], d["replicate_sweep"]["pt_tests_per_s"], "rep e2e", d["replicate_sweep"]["e2e"]["ms_per_step"])
    print("gemm", d["gemm_path"]["ms_per_step"], "cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r2_bench34.err
