#!/bin/bash
# round-2 GPU pass (2 GPUs): boundary chunks dealt over the CTAs, 8-way slot loads -- parity, 200k step times, 2M bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "test_two_ranks_halo_exchange_and_solve or env0" > gpurun_out/r2i_pytest_multi.log 2>&1
echo "pytest_multi rc=$?"; tail -3 gpurun_out/r2i_pytest_multi.log
for mode in ll nowait; do
  case $mode in
    ll) export NM_DEBUG_LL=0;;
    nowait) export NM_DEBUG_LL=1;;
  esac
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 60 --check-steps 0 --no-cpu > gpurun_out/r2i_bench_n2_$mode.json 2> gpurun_out/r2i_bench_n2_$mode.log
  echo "bench n2 $mode rc=$?"
  python - <<PY
import json
s=open('gpurun_out/r2i_bench_n2_$mode.json').read(); d=json.loads(s[s.index('{"metric'):])
print("$mode", {k: round(v['us'],2) for k,v in d['application']['kernels'].items()}, round(d['application']['us_per_degree_step'],1))
PY
done
unset NM_DEBUG_LL
time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2i_bench_2M_n2.json 2> gpurun_out/r2i_bench_2M_n2.log
echo "bench 2M n2 rc=$?"; grep -E "^\[bench\]" gpurun_out/r2i_bench_2M_n2.log | grep -v '^\[bench\] {' | tail -5
python - <<PY
import json
s=open('gpurun_out/r2i_bench_2M_n2.json').read(); d=json.loads(s[s.index('{"metric'):])
print("2M n2", {k: round(v['us'],2) for k,v in d['application']['kernels'].items()}, round(d['application']['us_per_degree_step'],1))
PY
