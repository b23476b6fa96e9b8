#!/bin/bash
mkdir -p gpurun_out
export DRNMF_REC_COOP=0
B="python bench.py --no-cpu-baseline --no-extras --no-throughput --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B --steps 2 --warmup 1 > gpurun_out/r2_under_ncu.json 2> gpurun_out/r2_prof.err
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_recurrent_tc -c 1 -o gpurun_out/r2_prof_recurrent $B --steps 1 --warmup 0 > /dev/null 2>> gpurun_out/r2_prof.err
echo "full recurrent rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 24 -c 1 -o gpurun_out/r2_prof_gemm $B --steps 1 --warmup 0 > /dev/null 2>> gpurun_out/r2_prof.err
echo "full gemm rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none -k regex:k_recurrent_tc -s 1 -c 1 --csv --log-file gpurun_out/r2_thr_metrics.csv python scripts/r2_prof_thr.py > gpurun_out/r2_thr.log 2>> gpurun_out/r2_prof.err
echo "thr metrics rc=$?"
tail -3 gpurun_out/r2_prof.err | cut -c1-200
grep -c k_ gpurun_out/r2_launches.csv; tail -8 gpurun_out/r2_thr_metrics.csv | cut -c1-260
