#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/trip3.log
timeout 900 python -m pytest tests -m gpu -x -q >> gpurun_out/trip3.log 2>&1
echo "rc=$?" >> gpurun_out/trip3.log
echo "=== bench" >> gpurun_out/trip3.log
timeout 900 python bench.py --steps 5 --warmup 3 --throughput-batch 512 > gpurun_out/bench_trip3.json 2>> gpurun_out/trip3.log
echo "rc=$?" >> gpurun_out/trip3.log
cat gpurun_out/bench_trip3.json >> gpurun_out/trip3.log
tail -40 gpurun_out/trip3.log
