#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/robust_sweep.py > gpurun_out/robust.log 2>&1
echo "rc=$?" >> gpurun_out/robust.log
timeout 600 python -m pytest tests -m gpu -x -q -k "full_size" >> gpurun_out/robust.log 2>&1
tail -30 gpurun_out/robust.log
