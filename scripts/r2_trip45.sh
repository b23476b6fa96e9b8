#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -x -k "split or ragged or throughput" ) > gpurun_out/r2t45_tests.log 2>&1; tail -3 gpurun_out/r2t45_tests.log | cut -c1-300
timeout 600 python scripts/r2_sweep.py ab 2>&1 | cut -c1-200
timeout 600 python scripts/r2_sweep.py rings 2>&1 | cut -c1-200 | grep "^B="
timeout 600 python scripts/r2_sweep.py dbg2 2>&1 | cut -c1-200 | grep -v "^B=" | head -12
