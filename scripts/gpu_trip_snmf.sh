#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "snmf" > gpurun_out/snmf.log 2>&1
echo "rc=$?" >> gpurun_out/snmf.log
tail -40 gpurun_out/snmf.log
