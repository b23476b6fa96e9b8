#!/bin/bash
# beta-divergence MU: parity tests + timing next to the Euclidean branch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "snmf" 2>&1 | tail -3
for b in 2 1 0 0.5; do MU_BETA=$b MU_FRAMES=22528 MU_ITERS=20 timeout 300 python scripts/mu_scaling.py; done 2>&1 | tee gpurun_out/t50_mu_beta.txt
for b in 2 1 0.5; do MU_BETA=$b MU_FRAMES=225000 MU_ITERS=10 timeout 300 python scripts/mu_scaling.py; done 2>&1 | tee -a gpurun_out/t50_mu_beta.txt
