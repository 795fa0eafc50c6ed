#!/bin/bash
# ncu capture of the final per-cell register kernel (10k cells x 50k SNVs)
mkdir -p gpurun_out
timeout 110 ncu --set full --clock-control none --import-source on -k regex:scs_place_percell -c 1 -o gpurun_out/r2f_prof_percell_final -f python bench.py --steps 1 --warmup 1 --snvs 50000 --no-gemm-leg --no-small-leg --replicates 0 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_percell_final.log 2>&1; echo "ncu rc=$?"
