#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "snmf or mu_ed or ista" 2>&1 | tail -2
for b in 2 1; do MU_BETA=$b MU_FRAMES=225000 MU_ITERS=10 timeout 300 python scripts/mu_scaling.py; done 2>&1 | cut -c1-330 | tee gpurun_out/t65_mu.txt
MU_FRAMES=22528 MU_ITERS=20 timeout 300 python scripts/mu_scaling.py 2>&1 | cut -c1-330 | tee -a gpurun_out/t65_mu.txt
