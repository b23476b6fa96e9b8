#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/r2_sweep.py llt ) 2>&1 | cut -c1-220
