#!/bin/bash
# 2 GPUs: data-parallel training line and the inference line with every pipeline on
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 600 $TR --master-port 29612 bench.py --workload train --gpus 2 --steps 8 --warmup 3 > gpurun_out/t72_train_2gpu.json 2> gpurun_out/t72.log
echo "train rc=$?"
timeout 600 $TR --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/t72_bench_2gpu.json 2>> gpurun_out/t72.log
echo "infer rc=$?"
python - <<'PY'
import json
for f in ('t72_train_2gpu','t72_bench_2gpu'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, {k:d.get(k) for k in ('n_gpus','value','ms_per_step')}, (d.get('e2e') or {}).get('value'))
PY
