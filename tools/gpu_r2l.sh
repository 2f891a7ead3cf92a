#!/bin/bash
# round-2 GPU check (2 GPUs, 200k mesh): 2-way slot loads (62 registers, 2 CTAs per SM) with / without dealt boundary chunks, pers
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 40 --check-steps 2 > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.log
  rc=$?
  python - <<PY
import json
try:
    s=open('gpurun_out/r2l_$name.json').read(); d=json.loads(s[s.index('{"metric'):])
    print("$name rc=$rc", {k: round(v['us'],2) for k,v in d['application']['kernels'].items()}, round(d['application']['us_per_degree_step'],1), "check", d['check']['max_rel_err'] if d.get('check') else None)
except Exception as e:
    print("$name rc=$rc failed", e)
PY
}
run nodeal NM_SLAB_DEAL_GHOST=0
run deal NM_SLAB_DEAL_GHOST=1
run pers_deal NM_SLAB_DEAL_GHOST=1 NM_SLAB_PERS=1 NM_SLAB_FLOW=0
