#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "forward or north or loss_and" > gpurun_out/l2.log 2>&1
echo "rc=$?" >> gpurun_out/l2.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_recurrent_tc -c 1 --csv --log-file gpurun_out/l2_metrics.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>> gpurun_out/l2.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l2.json 2>> gpurun_out/l2.log
tail -5 gpurun_out/l2.log; grep -v "^==" gpurun_out/l2_metrics.csv | cut -d, -f5,13- | head; python -c "
import json; d=json.load(open('gpurun_out/bench_l2.json')); print('value %.0f e2e %.0f'%(d['value'],d['e2e']['value']), d['roofline']['stage_ms'])"
