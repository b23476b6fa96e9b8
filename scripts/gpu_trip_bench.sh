#!/bin/bash
mkdir -p gpurun_out
echo "=== bench (default flags + extras)" > gpurun_out/bench_run.log
( time timeout 1200 python bench.py --throughput-batch 512 --extras > gpurun_out/bench_ours.json ) 2>> gpurun_out/bench_run.log
echo "rc=$?" >> gpurun_out/bench_run.log
echo "=== bench reference arm" >> gpurun_out/bench_run.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json ) 2>> gpurun_out/bench_run.log
echo "=== ncu launch list" >> gpurun_out/bench_run.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>> gpurun_out/bench_run.log
echo "rc=$?" >> gpurun_out/bench_run.log
echo "=== ncu full set on the recurrent kernel (full config)" >> gpurun_out/bench_run.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_recurrent_tc -c 1 -o gpurun_out/prof_recurrent_r1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>> gpurun_out/bench_run.log
echo "rc=$?" >> gpurun_out/bench_run.log
echo "=== ncu full set on the projection GEMM" >> gpurun_out/bench_run.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 24 -c 1 -o gpurun_out/prof_gemm_r1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>> gpurun_out/bench_run.log
echo "rc=$?" >> gpurun_out/bench_run.log
tail -30 gpurun_out/bench_run.log
