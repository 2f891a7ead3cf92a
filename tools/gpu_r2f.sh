#!/bin/bash
# round-2 sixth GPU pass (1 GPU): the driver's own commands on the 2 M-tet default workload, both arms
mkdir -p gpurun_out
free -g | head -2; nproc
time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.log
echo "bench n1 rc=$?"; grep -E "^\[bench\]" gpurun_out/r2f_bench_n1.log | grep -v '^\[bench\] {'; tail -3 gpurun_out/r2f_bench_n1.log
time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.log
echo "bench ref rc=$?"; tail -3 gpurun_out/r2f_bench_ref.log; cut -c1-300 gpurun_out/r2f_bench_ref.json
