#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/r2_sweep.py b64 ) > gpurun_out/r2t7_sweep.log 2> gpurun_out/r2t7_sweep.err
( timeout 600 python scripts/r2_sweep.py thr ) >> gpurun_out/r2t7_sweep.log 2>> gpurun_out/r2t7_sweep.err
( timeout 600 python scripts/r2_sweep.py trace ) > gpurun_out/r2t7_trace.log 2> gpurun_out/r2t7_trace.err
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x ) > gpurun_out/r2t7_parity.log 2>&1
cat gpurun_out/r2t7_sweep.log; tail -3 gpurun_out/r2t7_parity.log
