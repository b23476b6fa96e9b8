#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/r2_sweep.py ks2 2>&1 | cut -c1-230
