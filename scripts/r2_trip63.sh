#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "stft or istft or enhance or sdr or dataset or smoke" 2>&1 | tail -2
timeout 200 python scripts/stft_time.py
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:stft -c 8 --csv --log-file gpurun_out/t63_stft.csv python scripts/stft_time.py > /dev/null 2>&1
grep -v "^==" gpurun_out/t63_stft.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if r['ID'] in ('0','1','6','7'): print('  ', r['ID'], r['Kernel Name'][:34], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
"
