#!/bin/bash
# repeatability of the DRAM traffic of the B=64 recurrence launch (three captures on one box)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --no-throughput --no-parity"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for i in 1 2 3; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:k_recurrent_tc -s $i -c 1 --csv --log-file gpurun_out/t76_$i.csv $B --steps 3 --warmup 1 > /dev/null 2>&1
  grep -v "^==" gpurun_out/t76_$i.csv | python -c "
import csv,sys
print('capture $i:', ', '.join('%s %s' % (r['Metric Name'].split('__')[-1][:22], r['Metric Value']) for r in csv.DictReader(sys.stdin)))"
done
