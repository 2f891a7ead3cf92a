#!/bin/bash
# round-2 GPU pass (8 GPUs): the driver's own N=8 command on the 2 M-tet default workload
mkdir -p gpurun_out
free -g | head -2; nproc
time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n_bench_2M_n8.json 2> gpurun_out/r2n_bench_2M_n8.log
echo "bench 2M n8 rc=$?"; grep -E "^\[bench\]" gpurun_out/r2n_bench_2M_n8.log | grep -v '^\[bench\] {' | tail -12
python - <<PY
import json
try:
    s=open('gpurun_out/r2n_bench_2M_n8.json').read(); d=json.loads(s[s.index('{"metric'):])
    print("2M n8", {k: round(v['us'],2) for k,v in d['application']['kernels'].items()}, round(d['application']['us_per_degree_step'],1), d['check'], d['comm'])
except Exception as e: print("failed", e)
PY
tail -5 gpurun_out/r2n_bench_2M_n8.log | cut -c1-300
