#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/t71_pytest.log 2>&1
grep -E "passed|failed|Error" gpurun_out/t71_pytest.log | tail -3
timeout 600 python scripts/train_overlap_time.py 2>&1 | tee gpurun_out/t71_train_overlap.txt
timeout 600 python bench.py --workload train --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
t=json.loads(sys.stdin.read().strip()); print('train', t['value'], t['ms_per_step'], t['e2e']['value'])"
