#!/bin/bash
# A/B in one box: HEAD build vs working tree (debug accumulators moved to shared memory)
mkdir -p gpurun_out
for rep in 1 2; do
echo "== HEAD"; DRNMF_LIB=/root/repo/_head/dr-nmf_b200/libdrnmf.so timeout 600 python scripts/r2_sweep.py ab 2>&1 | cut -c1-200
echo "== TREE"; timeout 600 python scripts/r2_sweep.py ab 2>&1 | cut -c1-200
done
