#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/r2t46_tests.log 2>&1; tail -3 gpurun_out/r2t46_tests.log | cut -c1-300
bash scripts/r2_trip_final.sh
