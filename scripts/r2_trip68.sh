#!/bin/bash
# final verification of the round: full GPU suite, smoke, both bench arms, the training line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t68_pytest.log 2>&1
tail -4 gpurun_out/t68_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/t68_bench_ours.json ) 2> gpurun_out/t68_bench.log
echo "ours rc=$?"; grep real gpurun_out/t68_bench.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t68_bench_ref.json ) 2>> gpurun_out/t68_bench.log
echo "ref rc=$?"
timeout 600 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/t68_train_1gpu.json 2>> gpurun_out/t68_bench.log
echo "train rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t68_bench_ours.json').read().strip())
print('value %.0f e2e %.0f ms %.2f parity %s launches %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['ok'], d['gpu_launches']))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','pipe_frac_3xtf32','launch_ms','share_of_step')})
for k,v in d['config']['throughput_mode'].items(): print(k, {kk:v[kk] for kk in ('frames_per_s_forward_only','recurrence_ms','recurrence_useful_tflops','pipe_frac_3xtf32')})
for k,v in d['config']['extras'].items(): print(k, {kk:vv for kk,vv in v.items() if kk not in ('note','forward','backward')})
r=json.loads(open('gpurun_out/t68_bench_ref.json').read().strip()); print('ref', r['value'], r['cpu_baseline']['cores'])
t=json.loads(open('gpurun_out/t68_train_1gpu.json').read().strip()); print('train', t['value'], t['ms_per_step'], t['e2e']['value'])
PY
