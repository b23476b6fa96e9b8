#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python scripts/r2_stress.py ) > gpurun_out/r2t24_stress.log 2>&1; tail -16 gpurun_out/r2t24_stress.log | cut -c1-230
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu -x ) > gpurun_out/r2t24_tests.log 2>&1; tail -2 gpurun_out/r2t24_tests.log
