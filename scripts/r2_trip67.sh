#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "grads or trainer or pretrain or adam or train" 2>&1 | tail -3
timeout 600 python scripts/train_overlap_time.py 2>&1 | tee gpurun_out/t67_train_overlap.txt
