#!/bin/bash
# round-2 final GPU pass (1 GPU): the driver's GPU test command, smoke(), and a short bench on the 200k mesh after the cleanup
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest_gpu.log 2>&1
echo "pytest gpu rc=$?"; tail -3 gpurun_out/r2p_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2p_smoke.log
timeout 300 python bench.py --ntet 200000 --steps 2 --warmup 3 --degree-steps 100 --e2e-steps 1 > gpurun_out/r2p_bench_200k.json 2> gpurun_out/r2p_bench_200k.log
echo "bench rc=$?"; grep -E "check|device-resident" gpurun_out/r2p_bench_200k.log
