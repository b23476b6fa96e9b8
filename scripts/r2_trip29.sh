#!/bin/bash
( timeout 600 python scripts/r2_sweep.py b64x ) > gpurun_out/r2t29.log 2>&1
grep "rec \|FAILED" gpurun_out/r2t29.log | cut -c1-220
