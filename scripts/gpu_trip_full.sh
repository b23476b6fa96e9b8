#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/full.log
timeout 900 python -m pytest tests -m gpu -x -q >> gpurun_out/full.log 2>&1
echo "rc=$?" >> gpurun_out/full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/full.log 2>&1
echo "smoke rc=$?" >> gpurun_out/full.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2>> gpurun_out/full.log
python - <<'PY' >> gpurun_out/full.log
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d['roofline']['stage_ms'])
PY
tail -30 gpurun_out/full.log
