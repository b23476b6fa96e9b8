#!/bin/bash
# after the pipelined projection: repeatability, ncu launch list (serial fallback under the profiler), 2-GPU bench
mkdir -p gpurun_out
( STRESS_REPS=4 timeout 600 python scripts/r2_stress.py ) > gpurun_out/t60_stress.log 2>&1; tail -3 gpurun_out/t60_stress.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/t60_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-throughput --no-extras --no-parity --no-cpu-baseline > gpurun_out/t60_ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/t60_ncu_bench.log | cut -c1-300
grep -c "k_recurrent_tc" gpurun_out/t60_launches.csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/t60_bench_2gpu.json 2> gpurun_out/t60_bench_2gpu.log
echo "2gpu rc=$?"; cut -c1-400 gpurun_out/t60_bench_2gpu.json
