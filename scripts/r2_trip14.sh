#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu -x ) > gpurun_out/r2t14_tests.log 2>&1
tail -4 gpurun_out/r2t14_tests.log
( timeout 600 python scripts/r2_sweep.py final ) > gpurun_out/r2t14_final.log 2>&1
grep "rec " gpurun_out/r2t14_final.log
( DRNMF_REC_NOSYM=1 timeout 600 python scripts/r2_sweep.py b64 ) 2>&1 | grep "rec " | head -2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_recurrent_tc -c 3 --csv --log-file gpurun_out/r2t14_rec_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-throughput --no-parity > gpurun_out/r2t14_under_ncu.json 2> gpurun_out/r2t14_ncu.err
tail -8 gpurun_out/r2t14_rec_metrics.csv | cut -c1-300
