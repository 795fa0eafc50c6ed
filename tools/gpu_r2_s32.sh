#!/bin/bash
# final code of the round: whole GPU suite, smoke, the default bench line, one ncu capture of the per-cell kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest32.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest32.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke32.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke32.log
timeout 600 python bench.py > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench32.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "ms_per_launch", "prepass_ms")})
    print("per_cell_fn", d.get("per_cell_fn"))
    print("small", d["single_test_1k_cells"]["ms_per_step"], "rep", d["replicate_sweep"]["ms_per_step"], d["replicate_sweep"]["pt_tests_per_s"])
    print("gemm", d["gemm_path"]["ms_per_step"], "cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/r2_bench32.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scs_place_percell -c 1 -o gpurun_out/r2f_prof_percell -f python bench.py --steps 1 --warmup 1 --snvs 100000 --no-gemm-leg --no-small-leg --replicates 0 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_percell.log 2>&1; echo "ncu rc=$?"
