#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu ) > gpurun_out/r2t25_tests.log 2>&1
tail -25 gpurun_out/r2t25_tests.log | cut -c1-250
