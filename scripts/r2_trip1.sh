#!/bin/bash
# round 2, trip 1: parity at the benchmarked configs + operating-point sweep + per-role counters
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2t1_gpu.txt
( time timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "not 4000" ) > gpurun_out/r2t1_configs.log 2>&1
echo "rc=$?" >> gpurun_out/r2t1_configs.log
( time timeout 600 python scripts/r2_sweep.py all ) > gpurun_out/r2t1_sweep.log 2> gpurun_out/r2t1_sweep.err
echo "rc=$?" >> gpurun_out/r2t1_sweep.log
( DRNMF_REC_COOP=1 timeout 120 python scripts/rec_debug.py 64 20 ) > gpurun_out/r2t1_coop.log 2>&1
echo "rc=$?" >> gpurun_out/r2t1_coop.log
( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu ) > gpurun_out/r2t1_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r2t1_parity.log
tail -5 gpurun_out/r2t1_configs.log gpurun_out/r2t1_parity.log gpurun_out/r2t1_coop.log
tail -40 gpurun_out/r2t1_sweep.log
