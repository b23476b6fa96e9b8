#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/r2_sweep.py b64 ) > gpurun_out/r2t5_sweep.log 2> gpurun_out/r2t5_sweep.err
( timeout 600 python scripts/r2_sweep.py trace ) > gpurun_out/r2t5_trace.log 2> gpurun_out/r2t5_trace.err
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "forward_vs_oracle or golden or north_star" ) > gpurun_out/r2t5_parity.log 2>&1
cat gpurun_out/r2t5_sweep.log; tail -3 gpurun_out/r2t5_parity.log
