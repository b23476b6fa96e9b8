#!/bin/bash
mkdir -p gpurun_out
( SWEEP_K=7 timeout 600 python scripts/r2_sweep.py l2 ) > gpurun_out/r2t10_k7.log 2>&1
( SWEEP_K=13 timeout 600 python scripts/r2_sweep.py l2 ) > gpurun_out/r2t10_k13.log 2>&1
( SWEEP_K=25 DRNMF_REC_LL=0 timeout 600 python scripts/r2_sweep.py l2 ) > gpurun_out/r2t10_k25_noll.log 2>&1
( SWEEP_K=7 DRNMF_REC_LL=0 timeout 600 python scripts/r2_sweep.py l2 ) > gpurun_out/r2t10_k7_noll.log 2>&1
for f in k7 k13 k25_noll k7_noll; do echo "== $f"; grep "rec " gpurun_out/r2t10_$f.log; done
