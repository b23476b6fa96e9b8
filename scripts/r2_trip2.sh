#!/bin/bash
mkdir -p gpurun_out
( timeout 300 scripts/microbench/dsmem_xchg ) > gpurun_out/r2t2_dsmem.log 2>&1
echo "rc=$?" >> gpurun_out/r2t2_dsmem.log
( timeout 600 python scripts/r2_sweep.py ks4 ) > gpurun_out/r2t2_ks4.log 2> gpurun_out/r2t2_ks4.err
echo "rc=$?" >> gpurun_out/r2t2_ks4.log
cat gpurun_out/r2t2_dsmem.log gpurun_out/r2t2_ks4.log; grep libdrnmf gpurun_out/r2t2_ks4.err | sort | uniq | head -40
