#!/bin/bash
timeout 600 python scripts/r2_sweep.py rings2 2>&1 | cut -c1-200 | grep "^B="
