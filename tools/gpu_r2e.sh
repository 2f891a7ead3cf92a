#!/bin/bash
# round-2 fifth GPU pass (2 GPUs): whole single-GPU suite (P2 tests, RR refinement, lanbounds), N=2 step-time diagnostics
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py -k "not mtopo100k" -s --durations=8 > gpurun_out/r2e_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; grep -E "tight mode|eigenpairs,|passed|failed|Error|^[0-9.]+s " gpurun_out/r2e_pytest_gpu.log | tail -20
for mode in ll ll_nowait pers flags; do
  case $mode in
    ll) export NM_DEBUG_LL=0 NM_SLAB_PERS=0 NM_HALO_FUSED=1;;
    ll_nowait) export NM_DEBUG_LL=1 NM_SLAB_PERS=0 NM_HALO_FUSED=1;;
    pers) export NM_DEBUG_LL=0 NM_SLAB_PERS=1 NM_SLAB_FLOW=0 NM_HALO_FUSED=1;;
    flags) export NM_DEBUG_LL=0 NM_SLAB_PERS=0 NM_HALO_FUSED=0;;
  esac
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 100 --check-steps 0 --no-cpu > gpurun_out/r2e_bench_n2_$mode.json 2> gpurun_out/r2e_bench_n2_$mode.log
  echo "bench n2 $mode rc=$?"; grep -E "device-resident" gpurun_out/r2e_bench_n2_$mode.log
  python - <<PY
import json
s=open('gpurun_out/r2e_bench_n2_$mode.json').read(); d=json.loads(s[s.index('{"metric'):])
print("$mode", {k: round(v['us'],2) for k,v in d['application']['kernels'].items()}, round(d['application']['us_per_degree_step'],1))
PY
done
