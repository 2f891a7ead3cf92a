#!/bin/bash
# round-2 fourth GPU pass (2 GPUs): fused flag-in-data step as the multi-GPU default, Rayleigh-Ritz refinement, whole suite
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py -s > gpurun_out/r2d_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; grep -E "tight mode|passed|failed|Error" gpurun_out/r2d_pytest_gpu.log | tail -8
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2d_pytest_multi.log 2>&1
echo "pytest_multi rc=$?"; tail -5 gpurun_out/r2d_pytest_multi.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 0 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.log
echo "bench n2 rc=$?"; grep -E "check|device-resident|e2e|halo" gpurun_out/r2d_bench_n2.log
