#!/bin/bash
mkdir -p gpurun_out
for ft in 16 8 4; do
  DRNMF_STFT_FT=$ft timeout 600 ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"stft_mag_tiled|istft_ola_tiled" -s 8 -c 4 --csv --log-file gpurun_out/r2t26_ft$ft.csv python scripts/stft_time.py > /dev/null 2>&1
  echo "== FT=$ft"; python3 - <<PY
import csv
for r in csv.reader(open('gpurun_out/r2t26_ft$ft.csv')):
    if len(r)>14 and r[0].isdigit(): print(r[4][:40], r[12], r[14])
PY
done
