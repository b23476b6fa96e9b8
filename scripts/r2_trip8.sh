#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x ) > gpurun_out/r2t8_parity.log 2>&1
tail -3 gpurun_out/r2t8_parity.log
( timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu ) > gpurun_out/r2t8_configs.log 2>&1
tail -3 gpurun_out/r2t8_configs.log
( timeout 600 python scripts/r2_sweep.py b64 ) > gpurun_out/r2t8_sweep.log 2> gpurun_out/r2t8_sweep.err
( timeout 600 python scripts/r2_sweep.py b32 ) >> gpurun_out/r2t8_sweep.log 2>> gpurun_out/r2t8_sweep.err
( DRNMF_REC_LL=0 timeout 600 python scripts/r2_sweep.py b64 ) > gpurun_out/r2t8_sweep_noll.log 2> /dev/null
( timeout 600 python scripts/r2_sweep.py trace2 ) > gpurun_out/r2t8_trace.log 2> gpurun_out/r2t8_trace.err
cat gpurun_out/r2t8_sweep.log; head -3 gpurun_out/r2t8_sweep_noll.log
