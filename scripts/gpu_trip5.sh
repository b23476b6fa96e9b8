#!/bin/bash
mkdir -p gpurun_out
export DRNMF_REC_DEBUG=1
( timeout 300 python scripts/rec_debug.py 64 40
  timeout 300 python scripts/rec_debug.py 512 10 ) > gpurun_out/trip5.log 2>&1
tail -60 gpurun_out/trip5.log
