#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/r2_sweep.py trace2 ) > gpurun_out/r2t6_trace.log 2> gpurun_out/r2t6_trace.err
( DRNMF_REC_H2D=1 timeout 600 python scripts/r2_sweep.py trace2 ) > gpurun_out/r2t6_trace2d.log 2> gpurun_out/r2t6_trace2d.err
cat gpurun_out/r2t6_trace.log gpurun_out/r2t6_trace2d.log
