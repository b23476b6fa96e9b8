#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fwd_overlap_time.py 2>&1 | tee gpurun_out/t54_overlap.txt
timeout 1200 python -m pytest tests -m gpu -x -q -k "forward or north_star or edge or masked or enhance or hidden or golden" 2>&1 | tail -3
