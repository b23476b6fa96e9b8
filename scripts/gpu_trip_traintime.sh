#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/train_time.py 32 193 ) > gpurun_out/traintime.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/train_launches.csv python scripts/train_time.py 32 40 > /dev/null 2>> gpurun_out/traintime.log
tail -20 gpurun_out/traintime.log
