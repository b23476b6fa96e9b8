#!/bin/bash
# DRAM traffic and pipe utilisation of the B=64 recurrence launch of the final build (ncu, serial order under the profiler)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --no-throughput --no-parity"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread,sm__inst_executed.avg.per_cycle_active,smsp__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_recurrent_tc -s 1 -c 1 --csv --log-file gpurun_out/t75_b64_metrics.csv $B --steps 1 --warmup 1 > /dev/null 2> gpurun_out/t75.err
echo "rc=$?"
grep -v "^==" gpurun_out/t75_b64_metrics.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print('  ', r['Kernel Name'][:36], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
"
