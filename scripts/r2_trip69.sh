#!/bin/bash
# usage: gpurun --gpus 8 -- bash scripts/r2_trip69.sh   (8-GPU end points of the final build: inference, training, MU)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 600 $TR --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/t69_bench_8gpu.json 2> gpurun_out/t69.log
echo "infer rc=$?"
timeout 600 $TR --master-port 29612 bench.py --workload train --gpus 8 --steps 8 --warmup 3 > gpurun_out/t69_train_8gpu.json 2>> gpurun_out/t69.log
echo "train rc=$?"
MU_ITERS=40 timeout 600 $TR --master-port 29613 scripts/mu_scaling.py > gpurun_out/t69_mu_8gpu.json 2>> gpurun_out/t69.log
echo "mu rc=$?"
python - <<'PY'
import json
for f in ('t69_bench_8gpu','t69_train_8gpu','t69_mu_8gpu'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('n_gpus','value','ms_per_step','ms_per_iteration','useful_tflops_total')}, (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
