#!/bin/bash
timeout 600 python scripts/r2_sweep.py dbg2 2>&1 | cut -c1-200
