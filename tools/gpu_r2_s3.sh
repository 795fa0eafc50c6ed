#!/bin/bash
# round 2, session 2: tests, LRT barrier-parameter probe, bench, pre-pass comparison, ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest2.log | cut -c1-400
python bench.py --steps 5 --warmup 3 --no-gemm-leg > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
CMD="python bench.py --snvs 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --replicates 0 --no-gemm-leg"
$CMD > gpurun_out/r2_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:scs_" -c 200 --csv --log-file gpurun_out/r2_launches_s3_250k.csv $CMD > gpurun_out/r2_ncu_list2.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:scs_place_iv4|scs_permute_reduce" -s 2 -c 2 -o gpurun_out/r2_prof_s3_250k -f $CMD > gpurun_out/r2_ncu_full2.log 2>&1
echo "full capture rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench3.json", "gpurun_out/r2_bench3_prep1.json"):
    try:
        d = json.load(open(f))
        r = d["roofline"]
        print(f, "ms/step", d["ms_per_step"], "kernel", r["ms_per_launch"], "e2e", (d.get("e2e") or {}).get("ms_per_step"), "chunks", (d.get("e2e") or {}).get("chunks"))
    except Exception as e:
        print(f, "failed", e)
PY
