#!/bin/bash
# first GPU trip: SIMT end-to-end parity + tcgen05 GEMMs (with the SIMT recurrence)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/trip1_gpu.txt 2>&1
echo "=== simt only" > gpurun_out/trip1.log
DRNMF_RECURRENT=simt timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "simt or stft" >> gpurun_out/trip1.log 2>&1
echo "=== tc gemm + simt recurrent" >> gpurun_out/trip1.log
DRNMF_RECURRENT=simt timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tc" >> gpurun_out/trip1.log 2>&1
tail -40 gpurun_out/trip1.log
