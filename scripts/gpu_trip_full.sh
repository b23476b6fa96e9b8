#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/full.log
timeout 1200 python -m pytest tests -m gpu -x -q >> gpurun_out/full.log 2>&1
echo "rc=$?" >> gpurun_out/full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/full.log 2>&1
echo "smoke rc=$?" >> gpurun_out/full.log
tail -30 gpurun_out/full.log
