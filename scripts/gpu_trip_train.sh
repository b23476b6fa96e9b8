#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "loss_and_grads" > gpurun_out/train.log 2>&1
echo "rc=$?" >> gpurun_out/train.log
tail -60 gpurun_out/train.log
