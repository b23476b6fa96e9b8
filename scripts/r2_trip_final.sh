#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/r2_bench_ours.json ) 2> gpurun_out/r2_final.log
echo "ours rc=$?" >> gpurun_out/r2_final.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json ) 2>> gpurun_out/r2_final.log
echo "ref rc=$?" >> gpurun_out/r2_final.log
timeout 600 python bench.py --workload train --steps 8 --warmup 3 > gpurun_out/r2_train_1gpu.json 2>> gpurun_out/r2_final.log
timeout 600 python bench.py --workload train --seconds 7.9 --steps 4 --warmup 2 > gpurun_out/r2_train_1gpu_T500.json 2>> gpurun_out/r2_final.log
echo "train rc=$?" >> gpurun_out/r2_final.log
( timeout 600 python scripts/r2_sweep.py dbg ) > gpurun_out/r2_cycle_table.log 2> gpurun_out/r2_cycle_table.txt
( timeout 600 python scripts/r2_sweep.py final ) > gpurun_out/r2_recurrence_sweep.txt 2>&1
( timeout 600 python scripts/r2_stress.py ) > gpurun_out/r2_stress.log 2>&1; tail -1 gpurun_out/r2_stress.log
grep "rc=\|real" gpurun_out/r2_final.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_ours.json').read().strip())
print('value %.0f e2e %.0f ms %.2f parity %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','pipe_frac_3xtf32','launch_ms','share_of_step','traffic')})
for k,v in d['config']['throughput_mode'].items(): print(k, {kk:v[kk] for kk in ('frames_per_s_forward_only','recurrence_ms','recurrence_useful_tflops','pipe_frac_3xtf32')})
for k,v in d['config']['extras'].items(): print(k, {kk:vv for kk,vv in v.items() if kk not in ('note','forward','backward')})
r=json.loads(open('gpurun_out/r2_bench_ref.json').read().strip()); print('ref', r['value'], r['cpu_baseline']['cores'])
for f in ('r2_train_1gpu','r2_train_1gpu_T500'):
    t=json.loads(open('gpurun_out/%s.json'%f).read().strip()); print(f, t['value'], t['ms_per_step'], t['config']['T'], t['e2e']['value'])
PY
