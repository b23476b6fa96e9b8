#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "loss_and_grads or fit" > gpurun_out/train.log 2>&1
echo "rc=$?" >> gpurun_out/train.log
timeout 600 python scripts/train_time.py 32 193 >> gpurun_out/train.log 2>&1
tail -30 gpurun_out/train.log
