#!/bin/bash
# usage: gpurun --gpus 8 -- bash scripts/r2_trip73.sh   (8-GPU data-parallel training line of the final build)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29612 bench.py --workload train --gpus 8 --steps 8 --warmup 3 > gpurun_out/t73_train_8gpu.json 2> gpurun_out/t73.log
echo "train rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/t73_train_8gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step')}, d['e2e']['value'])"
