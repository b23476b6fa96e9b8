#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/prec.log 2>&1
echo "rc=$?" >> gpurun_out/prec.log
timeout 1500 python scripts/precision_check.py 8 >> gpurun_out/prec.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prec.json 2>> gpurun_out/prec.log
python -c "
import json; d=json.load(open('gpurun_out/bench_prec.json')); print('value %.0f e2e %.0f'%(d['value'],d['e2e']['value']), d['roofline']['stage_ms'])" >> gpurun_out/prec.log
tail -22 gpurun_out/prec.log
