#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu ) > gpurun_out/r2t4_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r2t4_parity.log
tail -12 gpurun_out/r2t4_parity.log
( timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu ) > gpurun_out/r2t4_configs.log 2>&1
echo "rc=$?" >> gpurun_out/r2t4_configs.log
tail -8 gpurun_out/r2t4_configs.log
( timeout 600 python scripts/r2_sweep.py b64 ) > gpurun_out/r2t4_sweep.log 2> gpurun_out/r2t4_sweep.err
( timeout 600 python scripts/r2_sweep.py thr ) >> gpurun_out/r2t4_sweep.log 2>> gpurun_out/r2t4_sweep.err
( timeout 600 python scripts/r2_sweep.py trace ) > gpurun_out/r2t4_trace.log 2> gpurun_out/r2t4_trace.err
cat gpurun_out/r2t4_sweep.log
