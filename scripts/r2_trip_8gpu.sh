#!/bin/bash
# usage: gpurun --gpus 8 -- bash scripts/r2_trip_8gpu.sh
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_8gpu.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then L="python"; else L="$TR --nproc-per-node $N --master-port $((29500+N))"; fi
  timeout 600 $L bench.py --workload train --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_train_${N}gpu.json 2>> gpurun_out/r2_8gpu.log
  echo "train N=$N rc=$?" >> gpurun_out/r2_8gpu.log
  MU_ITERS=40 timeout 600 $L scripts/mu_scaling.py > gpurun_out/r2_mu_${N}gpu.json 2>> gpurun_out/r2_8gpu.log
  echo "mu N=$N rc=$?" >> gpurun_out/r2_8gpu.log
done
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2>> gpurun_out/r2_8gpu.log
echo "infer N=8 rc=$?" >> gpurun_out/r2_8gpu.log
for R in 1000 2000 4000; do
  timeout 900 $TR --nproc-per-node 8 --master-port $((29700+R/1000)) bench.py --gpus 8 --R $R --nfft 2048 --hop 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_sweep_R${R}_F1025_8gpu.json 2>> gpurun_out/r2_8gpu.log
  echo "sweep R=$R rc=$?" >> gpurun_out/r2_8gpu.log
done
timeout 300 $TR --nproc-per-node 8 --master-port 29811 tests/dist_snmf_check.py 2>&1 | grep dist_snmf >> gpurun_out/r2_8gpu.log
grep "rc=\|dist_snmf" gpurun_out/r2_8gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_*gpu.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('n_gpus','value','ms_per_step','ms_per_iteration','useful_tflops_total')}, (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
