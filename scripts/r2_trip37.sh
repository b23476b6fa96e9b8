#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2t37_tests.log 2>&1; tail -3 gpurun_out/r2t37_tests.log | cut -c1-300
( timeout 600 python scripts/r2_stress.py ) > gpurun_out/r2_stress.log 2>&1; tail -1 gpurun_out/r2_stress.log
( timeout 600 python scripts/r2_sweep.py final ) > gpurun_out/r2_recurrence_sweep.txt 2>&1
cat gpurun_out/r2_recurrence_sweep.txt | cut -c1-220
( timeout 600 python scripts/r2_sweep.py dbg ) > gpurun_out/r2_cycle_table.log 2> gpurun_out/r2_cycle_table.txt
cat gpurun_out/r2_cycle_table.txt | cut -c1-200
