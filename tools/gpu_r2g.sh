#!/bin/bash
# round-2 diagnostic (2 GPUs): what bounds the multi-GPU ChebIter step -- LL tag waits, LL peer stores, or neither
mkdir -p gpurun_out
for mode in ll nowait nowait_nostore; do
  case $mode in
    ll) export NM_DEBUG_LL=0;;
    nowait) export NM_DEBUG_LL=1;;
    nostore) export NM_DEBUG_LL=2;;
    nowait_nostore) export NM_DEBUG_LL=3;;
  esac
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 60 --check-steps 0 --no-cpu > gpurun_out/r2g_bench_n2_$mode.json 2> gpurun_out/r2g_bench_n2_$mode.log
  echo "bench n2 $mode rc=$?"
  python - <<PY
import json
s=open('gpurun_out/r2g_bench_n2_$mode.json').read(); d=json.loads(s[s.index('{"metric'):])
print("$mode", {k: round(v['us'],2) for k,v in d['application']['kernels'].items()}, round(d['application']['us_per_degree_step'],1))
PY
done
