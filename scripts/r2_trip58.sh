#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fwd_overlap_time.py 2>&1 | tee gpurun_out/t58_overlap.txt
CUDA_LAUNCH_BLOCKING=1 DRNMF_FWD_OVERLAP=force timeout 300 python scripts/fwd_overlap_time.py blocking 2>&1 | tail -3 | tee -a gpurun_out/t58_overlap.txt
CUDA_LAUNCH_BLOCKING=1 timeout 300 python scripts/fwd_overlap_time.py blocking 2>&1 | tail -2 | tee -a gpurun_out/t58_overlap.txt
