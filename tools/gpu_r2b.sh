#!/bin/bash
# round-2 second GPU pass (2 GPUs): persistent ChebIter kernel -- single-GPU parity, multi-GPU parity, kernel times, N=2 bench
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "slab_kernel_configs or jacobi or solve_prem3k or filter_application" > gpurun_out/r2b_pytest_pers.log 2>&1
echo "pytest pers rc=$?"; tail -15 gpurun_out/r2b_pytest_pers.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2b_pytest_multi.log 2>&1
echo "pytest_multi rc=$?"; tail -15 gpurun_out/r2b_pytest_multi.log
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/kernel_times.py --out gpurun_out/r2b_kernel_times.json > gpurun_out/r2b_kernel_times.log 2>&1
echo "kernel_times rc=$?"; grep -E "us " gpurun_out/r2b_kernel_times.log | tail -12
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.log
echo "bench n2 rc=$?"; grep -E "device-resident|e2e|halo" gpurun_out/r2b_bench_n2.log
