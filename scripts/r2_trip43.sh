#!/bin/bash
timeout 600 python scripts/r2_sweep.py rings 2>&1 | cut -c1-200 | grep "^B="
timeout 600 python scripts/r2_sweep.py rings 2>&1 | cut -c1-200 | grep "^B="
