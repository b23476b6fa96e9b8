#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "stft or istft or enhance or sdr or dataset or smoke" 2>&1 | tail -3
echo "in-place:"; timeout 200 python scripts/stft_time.py
echo "staged  :"; DRNMF_ISTFT_STAGED=1 timeout 200 python scripts/stft_time.py
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:istft -c 2 --csv --log-file gpurun_out/t62_istft.csv python scripts/stft_time.py > /dev/null 2>&1
grep -v "^==" gpurun_out/t62_istft.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print('  ', r['Kernel Name'][:34], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
"
