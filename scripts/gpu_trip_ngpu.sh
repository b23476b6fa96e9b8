#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_trip_ngpu.sh N   (weak-scaling bench line at N GPUs + the 2-GPU sharded-MU check)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/ngpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2>> gpurun_out/ngpu.log
echo "rc=$?" >> gpurun_out/ngpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_snmf_check.py 2>&1 | grep dist_snmf >> gpurun_out/ngpu.log
tail -6 gpurun_out/ngpu.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print('n_gpus',d['n_gpus'],'value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d.get('clocks'))
PY
