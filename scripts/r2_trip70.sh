#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "enhance or pipelined or forward or smoke or sdr" 2>&1 | tail -2
for m in 1 0; do
  DRNMF_FWD_OVERLAP=$m timeout 600 python bench.py --no-throughput --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/t70_bench_ov$m.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/t70_bench_ov$m.json').read().strip())
print('overlap=$m value %.0f (%.2f ms)  e2e %.0f (%.2f ms)  parity %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity']['ok'] if d.get('parity') else None))
PY
done
