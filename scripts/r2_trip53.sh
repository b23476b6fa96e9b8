#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/r2_sweep.py ab 2>&1 | tee gpurun_out/t53_ab.txt
timeout 1200 python -m pytest tests -m gpu -x -q -k "forward or grads or north_star or edge or masked" 2>&1 | tail -3
