#!/bin/bash
# round-2 GPU diagnostic (2 GPUs, 200k mesh): cost of the flag-in-data slot accesses by load / store flavour
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 40 --check-steps 2 > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.log
  rc=$?
  python - <<PY
import json
try:
    s=open('gpurun_out/r2j_$name.json').read(); d=json.loads(s[s.index('{"metric'):])
    print("$name rc=$rc", {k: round(v['us'],2) for k,v in d['application']['kernels'].items()}, round(d['application']['us_per_degree_step'],1), "check", d['check']['max_rel_err'] if d.get('check') else None)
except Exception as e:
    print("$name rc=$rc failed", e)
PY
}
run vol_vol NM_DEBUG_LL=0
run cg_vol NM_DEBUG_LL=4
run gpu_vol NM_DEBUG_LL=8
run vol_cg NM_DEBUG_LL=16
run cg_cg NM_DEBUG_LL=20
run cg_cg_deal NM_DEBUG_LL=20 NM_SLAB_DEAL_GHOST=1
run flags NM_HALO_FUSED=0
