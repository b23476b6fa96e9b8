#!/bin/bash
# tcgen05 persistent recurrence: parity
mkdir -p gpurun_out
echo "=== tc full" > gpurun_out/trip2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc" >> gpurun_out/trip2.log 2>&1
echo "rc=$?" >> gpurun_out/trip2.log
tail -60 gpurun_out/trip2.log
