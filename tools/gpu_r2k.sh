#!/bin/bash
# round-2 GPU diagnostic (2 GPUs, 200k mesh): ncu --set full of the fused multi-GPU ChebIter step on rank 0 (rank 1 runs plain)
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29711 WORLD_SIZE=2 NM_DEBUG_LL=1
ARGS="bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 --ntet 200000 --degree-steps 4 --check-steps 0 --no-cpu"
RANK=1 LOCAL_RANK=1 timeout 600 python $ARGS > gpurun_out/r2k_rank1.log 2>&1 &
P1=$!
RANK=0 LOCAL_RANK=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_slabws --launch-skip 41 -c 10 -f -o gpurun_out/r2k_slabws_n2 python $ARGS > gpurun_out/r2k_rank0.log 2>&1
echo "rank0 rc=$?"
wait $P1; echo "rank1 rc=$?"
tail -3 gpurun_out/r2k_rank0.log
ncu -i gpurun_out/r2k_slabws_n2.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread > gpurun_out/r2k_slabws_n2_summary.csv 2>&1
cut -c1-400 gpurun_out/r2k_slabws_n2_summary.csv | tail -14
