#!/bin/bash
# final-build profiles: launch list of the bench step, full-set metrics of the B=64 recurrence launch (traffic) and the
# throughput-mode (split epilogue) launch
mkdir -p gpurun_out
export DRNMF_REC_COOP=0
B="python bench.py --no-cpu-baseline --no-extras --no-throughput --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B --steps 2 --warmup 1 > gpurun_out/r2_under_ncu.json 2> gpurun_out/r2_prof.err
echo "launch list rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread,sm__inst_executed.avg.per_cycle_active,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:k_recurrent_tc -s 1 -c 1 --csv --log-file gpurun_out/r2_b64_metrics.csv $B --steps 1 --warmup 1 > /dev/null 2>> gpurun_out/r2_prof.err
echo "b64 metrics rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:k_recurrent_tc -s 1 -c 1 --csv --log-file gpurun_out/r2_thr_metrics.csv python scripts/r2_prof_thr.py > gpurun_out/r2_thr.log 2>> gpurun_out/r2_prof.err
echo "thr metrics rc=$?"
PROF_B=512 PROF_T=48 timeout 900 ncu --metrics $M --clock-control none -k regex:k_recurrent_tc -s 1 -c 1 --csv --log-file gpurun_out/r2_thr512_metrics.csv python scripts/r2_prof_thr.py > gpurun_out/r2_thr512.log 2>> gpurun_out/r2_prof.err
echo "thr512 metrics rc=$?"
tail -3 gpurun_out/r2_prof.err | cut -c1-200
grep -c k_ gpurun_out/r2_launches.csv
for f in r2_b64_metrics r2_thr_metrics r2_thr512_metrics; do echo "== $f"; grep -v "^==" gpurun_out/$f.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print('  ', r['Kernel Name'][:40], r['Grid Size'], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
"; done
