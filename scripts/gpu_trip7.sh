#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/trip7.log
timeout 900 python -m pytest tests -m gpu -x -q >> gpurun_out/trip7.log 2>&1
echo "rc=$?" >> gpurun_out/trip7.log
export DRNMF_REC_DEBUG=1
( timeout 300 python scripts/rec_debug.py 64 40
  DRNMF_REC_NB=32 timeout 300 python scripts/rec_debug.py 64 40
  timeout 300 python scripts/rec_debug.py 512 10 ) 2>&1 | grep -v "^\[libdrnmf\] recurrence debug" | awk '/h-loader/{c++} { if (c%2==0 || /^B=/) print }' >> gpurun_out/trip7.log 2>&1
tail -40 gpurun_out/trip7.log
