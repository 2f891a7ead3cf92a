#!/bin/bash
# round-2 third GPU pass (2 GPUs): dataflow persistent kernel -- parity on 1 and 2 GPUs, kernel times, N=2 bench on the 200k mesh
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "slab_kernel_configs or jacobi or solve_prem3k or filter_application" > gpurun_out/r2c_pytest_pers.log 2>&1
echo "pytest pers rc=$?"; tail -5 gpurun_out/r2c_pytest_pers.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2c_pytest_multi.log 2>&1
echo "pytest_multi rc=$?"; tail -5 gpurun_out/r2c_pytest_multi.log
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/kernel_times.py --out gpurun_out/r2c_kernel_times_flow.json > gpurun_out/r2c_kernel_times_flow.log 2>&1
echo "kernel_times flow rc=$?"; grep -E "chebiter_step" gpurun_out/r2c_kernel_times_flow.log
CUDA_VISIBLE_DEVICES=0 NM_SLAB_PERS_STAGES=2 timeout 400 python tools/kernel_times.py --out gpurun_out/r2c_kernel_times_flow_s2.json > gpurun_out/r2c_kernel_times_flow_s2.log 2>&1
echo "kernel_times flow stages=2 rc=$?"; grep -E "chebiter_step" gpurun_out/r2c_kernel_times_flow_s2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 0 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.log
echo "bench n2 rc=$?"; grep -E "check|device-resident|e2e|halo" gpurun_out/r2c_bench_n2.log
