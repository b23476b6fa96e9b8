#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/r2_sweep.py final ) > gpurun_out/r2t12_final.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu ) > gpurun_out/r2t12_tests.log 2>&1
grep "rec \|FAILED\|bitwise" gpurun_out/r2t12_final.log; tail -5 gpurun_out/r2t12_tests.log
