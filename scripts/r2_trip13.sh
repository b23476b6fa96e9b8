#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -k "trainer or device_error" ) > gpurun_out/r2t13_tests.log 2>&1
tail -15 gpurun_out/r2t13_tests.log
( time timeout 1200 python bench.py > gpurun_out/r2t13_bench.json ) 2> gpurun_out/r2t13_bench.err
tail -5 gpurun_out/r2t13_bench.err
( time timeout 600 python bench.py --workload train --steps 5 --warmup 2 > gpurun_out/r2t13_train.json ) 2> gpurun_out/r2t13_train.err
tail -5 gpurun_out/r2t13_train.err
cat gpurun_out/r2t13_train.json | cut -c1-1500
