#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/r2_sweep.py b64 ) > gpurun_out/r2t9_sweep.log 2> gpurun_out/r2t9_sweep.err
( timeout 600 python scripts/r2_sweep.py b32 ) >> gpurun_out/r2t9_sweep.log 2>> gpurun_out/r2t9_sweep.err
( timeout 600 python scripts/r2_sweep.py trace2 ) > gpurun_out/r2t9_trace.log 2> gpurun_out/r2t9_trace.err
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x ) > gpurun_out/r2t9_parity.log 2>&1
cat gpurun_out/r2t9_sweep.log; tail -2 gpurun_out/r2t9_parity.log
