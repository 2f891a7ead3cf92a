#!/bin/bash
# round-2 GPU pass (1 GPU): time-to-all-eigenpairs on the 2 M-tet mesh at P1; ncu of the DMMA kernels
mkdir -p gpurun_out
timeout 1500 python bench.py --porder 1 --solve --steps 2 --warmup 1 --degree-steps 64 --e2e-steps 1 --cpu-seconds 3 > gpurun_out/r2o_solve_2M_p1.json 2> gpurun_out/r2o_solve_2M_p1.log
echo "solve rc=$?"; grep -E "^\[bench\]" gpurun_out/r2o_solve_2M_p1.log | grep -v '^\[bench\] {' | tail -6
python - <<PY
import json
try:
    s=open('gpurun_out/r2o_solve_2M_p1.json').read(); d=json.loads(s[s.index('{"metric'):])
    print("solve", d.get('time_to_all_eigenpairs_s'), d.get('solve'), d['detail']['filter_degree'], d['application']['us_per_degree_step'])
except Exception as e: print("solve failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ritz_gemm|k_gram_dmma" -c 4 -f -o gpurun_out/r2o_dmma_kernels python tools/lanczos_kernels.py --n 814323 --k 200 --ns 64 --out gpurun_out/r2o_lanczos_kernels_ncu.json > gpurun_out/r2o_ncu_dmma.log 2>&1
echo "ncu dmma rc=$?"
