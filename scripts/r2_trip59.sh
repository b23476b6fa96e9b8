#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t59_pytest.log 2>&1
tail -4 gpurun_out/t59_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/t59_bench.json ) 2> gpurun_out/t59_bench.log
tail -4 gpurun_out/t59_bench.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t59_bench.json').read().strip())
print('value %.0f e2e %.0f ms %.2f parity_ok %s launches %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['ok'], d.get('gpu_launches')))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','pipe_frac_3xtf32','launch_ms','share_of_step','traffic')})
PY
