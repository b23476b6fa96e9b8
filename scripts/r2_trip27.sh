#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stft or sdr or smoke or golden" ) 2>&1 | tail -3
for ft in 16 8; do
  DRNMF_STFT_FT=$ft timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"stft_mag_tiled|istft_ola_tiled" -s 6 -c 2 --csv --log-file gpurun_out/r2t27_ft$ft.csv python scripts/stft_time.py > /dev/null 2>&1
  echo "== FT=$ft"; python3 - <<PY
import csv
for r in csv.reader(open('gpurun_out/r2t27_ft$ft.csv')):
    if len(r)>14 and r[0].isdigit(): print(r[4][:40], r[12], r[14])
PY
done
