#!/bin/bash
# round 2, trip 3: first run of the re-tiled recurrence (KS=4 clusters, batch groups, TMEM weight ring)
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2t3_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r2t3_smoke.log
tail -3 gpurun_out/r2t3_smoke.log
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu ) > gpurun_out/r2t3_parity.log 2>&1
echo "rc=$?" >> gpurun_out/r2t3_parity.log
tail -15 gpurun_out/r2t3_parity.log
( timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu ) > gpurun_out/r2t3_configs.log 2>&1
echo "rc=$?" >> gpurun_out/r2t3_configs.log
tail -15 gpurun_out/r2t3_configs.log
( timeout 600 python scripts/r2_sweep.py all ) > gpurun_out/r2t3_sweep.log 2> gpurun_out/r2t3_sweep.err
echo "rc=$?" >> gpurun_out/r2t3_sweep.log
cat gpurun_out/r2t3_sweep.log
