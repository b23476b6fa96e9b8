#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/2gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2>> gpurun_out/2gpu.log
echo "rc=$?" >> gpurun_out/2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-utts 8 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/2gpu.log
echo "rc=$?" >> gpurun_out/2gpu.log
tail -5 gpurun_out/2gpu.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print('n_gpus',d['n_gpus'],'value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))
print(open('gpurun_out/bench_2gpu_ref.json').read()[:300])
PY
