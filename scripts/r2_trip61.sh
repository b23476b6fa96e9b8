#!/bin/bash
# the profiling recipe as written (no library-specific environment): launch list of the bench step under ncu
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t61_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-throughput --no-extras --no-parity --no-cpu-baseline > gpurun_out/t61_ncu_bench.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/t61_ncu_bench.log | cut -c1-200
grep -c "k_recurrent_tc" gpurun_out/t61_launches.csv
python scripts/summarize_profiles.py gpurun_out/t61_launches.csv 2>/dev/null | head -20
