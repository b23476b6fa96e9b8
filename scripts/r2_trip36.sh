#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/r2_sweep.py ab 2>&1 | cut -c1-260
( timeout 600 python scripts/r2_stress.py ) > gpurun_out/r2_stress.log 2>&1; tail -2 gpurun_out/r2_stress.log | cut -c1-300
