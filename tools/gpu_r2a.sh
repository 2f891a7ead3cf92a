#!/bin/bash
# round-2 first GPU pass (2 GPUs): fused halo tests, fused N=2 bench line, single-GPU kernel times
mkdir -p gpurun_out
NM_TEST_FUSED=1 timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2a_pytest_multi.log 2>&1
echo "pytest_multi rc=$?"
tail -5 gpurun_out/r2a_pytest_multi.log
NM_HALO_FUSED=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu > gpurun_out/r2a_bench_n2_fused.json 2> gpurun_out/r2a_bench_n2_fused.log
echo "bench fused rc=$?"; cat gpurun_out/r2a_bench_n2_fused.json | cut -c1-600
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/kernel_times.py --out gpurun_out/r2a_kernel_times.json > gpurun_out/r2a_kernel_times.log 2>&1
echo "kernel_times rc=$?"; grep -E "us " gpurun_out/r2a_kernel_times.log | tail -12
