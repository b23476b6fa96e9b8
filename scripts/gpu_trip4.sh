#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/trip4.log
timeout 900 python -m pytest tests -m gpu -x -q >> gpurun_out/trip4.log 2>&1
echo "rc=$?" >> gpurun_out/trip4.log
echo "=== bench" >> gpurun_out/trip4.log
timeout 900 python bench.py --steps 5 --warmup 3 --throughput-batch 512 --no-cpu-baseline > gpurun_out/bench_trip4.json 2>> gpurun_out/trip4.log
echo "rc=$?" >> gpurun_out/trip4.log
python - <<'PY' >> gpurun_out/trip4.log
import json
d=json.load(open('gpurun_out/bench_trip4.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
print('stage',d['roofline']['stage_ms'])
print('rec',d['config']['recurrence'])
print('thr',d['config'].get('throughput_mode'))
PY
tail -30 gpurun_out/trip4.log
