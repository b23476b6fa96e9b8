#!/bin/bash
timeout 600 python scripts/r2_sweep.py ab 2>&1 | cut -c1-200
timeout 600 python scripts/r2_sweep.py rings 2>&1 | cut -c1-200
timeout 600 python scripts/r2_sweep.py dbg2 2>&1 | cut -c1-200 | grep -v "^B="
