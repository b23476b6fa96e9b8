#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/r2_sweep.py llt ) > gpurun_out/r2t11_llt.log 2>&1
( timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/r2_sweep.py crash ) > gpurun_out/r2t11_san.log 2>&1
grep "rec \|FAILED\|bitwise" gpurun_out/r2t11_llt.log; grep -v "^B=" gpurun_out/r2t11_san.log | head -40
