#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/r2_sweep.py ab 2>&1 | cut -c1-200
( timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2t40_tests.log 2>&1; tail -3 gpurun_out/r2t40_tests.log | cut -c1-300
