#!/bin/bash
# round-2 GPU pass (1 GPU): whole GPU suite, ncu launch list + full captures (2M-tet workload), Lanczos dense kernels, one full solve
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=6 > gpurun_out/r2m_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; grep -E "tight mode|eigenpairs,|passed|failed|rror" gpurun_out/r2m_pytest_gpu.log | tail -12
# time-to-all-eigenpairs: 200k-tet mesh, P1, JOB 2, 0.1-1.0 mHz
timeout 900 python bench.py --ntet 200000 --porder 1 --solve --steps 2 --warmup 1 --degree-steps 0 --e2e-steps 1 > gpurun_out/r2m_solve_200k_p1.json 2> gpurun_out/r2m_solve_200k_p1.log
echo "solve rc=$?"; python - <<PY
import json
try:
    s=open('gpurun_out/r2m_solve_200k_p1.json').read(); d=json.loads(s[s.index('{"metric'):])
    print("solve", d.get('time_to_all_eigenpairs_s'), d.get('solve'), d['detail']['filter_degree'], d['ms_per_step'])
except Exception as e: print("solve failed", e)
PY
timeout 300 python tools/lanczos_kernels.py --n 8364411 --k 200 500 --ns 128 --out gpurun_out/r2m_lanczos_kernels.json > gpurun_out/r2m_lanczos_kernels.log 2>&1
echo "lanczos kernels rc=$?"; tail -2 gpurun_out/r2m_lanczos_kernels.log | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_gemvT|k_gemvN|k_ritz_gemm|k_gram_dmma" --launch-skip 8 -c 6 -f -o gpurun_out/r2m_lanczos_kernels python tools/lanczos_kernels.py --n 814323 --k 200 --ns 64 --out gpurun_out/r2m_lanczos_kernels_ncu.json > gpurun_out/r2m_ncu_lanczos.log 2>&1
echo "ncu lanczos rc=$?"
# 2M-tet workload under ncu: launch list of one degree step, then --set full of 2 B~ and 2 Ap~ steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 160 --csv --log-file gpurun_out/r2m_launches_2M.csv python bench.py --steps 1 --warmup 1 --degree-steps 1 --check-steps 0 --no-cpu --e2e-steps 1 > gpurun_out/r2m_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r2m_launches_2M.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_slabws --launch-skip 43 -c 4 -f -o gpurun_out/r2m_kslabws_2M python bench.py --steps 1 --warmup 1 --degree-steps 1 --check-steps 0 --no-cpu --e2e-steps 1 > gpurun_out/r2m_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
